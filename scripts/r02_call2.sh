#!/bin/bash
mkdir -p gpurun_out
B=./cugraph_b200/lib/cbench
TAG=${1:-v2}
{
echo "== new sweep $TAG"; timeout 120 $B 24 sweep
echo "== pagerank"; timeout 120 $B 24 pagerank
} 2>&1 | tee gpurun_out/r02_cbench_$TAG.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 6 -c 2 -f -o gpurun_out/r02_ncu_sweep_$TAG $B 24 sweep > gpurun_out/r02_ncu_sweep_$TAG.log 2>&1
tail -2 gpurun_out/r02_ncu_sweep_$TAG.log
