#!/bin/bash
# A/B of two library builds (cugraph_b200/lib = default, cugraph_b200/lib_b = alternative), interleaved
mkdir -p gpurun_out
for r in 1 2; do
for L in lib lib_b; do echo "== $L"; timeout 120 ./cugraph_b200/$L/cbench 24 sweep; done
done 2>&1 | tee gpurun_out/r02_ab_${1:-x}.log
