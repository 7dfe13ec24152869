#!/bin/bash
# A/B of library builds (cugraph_b200/lib = default, cugraph_b200/lib_X = alternatives), interleaved
mkdir -p gpurun_out
VARS=${VARIANTS:-"lib lib_b"}
for r in 1 2; do
for L in $VARS; do echo "== $L"; CUGRAPH_B200_BUILD_TRACE=${TRACE:-0} timeout 120 ./cugraph_b200/$L/cbench 24 sweep; done
done 2>&1 | tee gpurun_out/r02_ab_${1:-x}.log
