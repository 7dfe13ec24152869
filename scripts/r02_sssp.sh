#!/bin/bash
mkdir -p gpurun_out
B=./cugraph_b200/lib/cbench
for cfg in "CUGRAPH_B200_SSSP_START_DIV=64" "CUGRAPH_B200_SSSP_START_DIV=16" "CUGRAPH_B200_SSSP_START_DIV=8" "CUGRAPH_B200_SSSP_START_DIV=4" "CUGRAPH_B200_SSSP_START_DIV=2" "CUGRAPH_B200_SSSP_START_DIV=8 CUGRAPH_B200_SSSP_SPLIT_ROUNDS=2" "CUGRAPH_B200_SSSP_START_DIV=16 CUGRAPH_B200_SSSP_SPLIT_MIN_EDGES=100000000000"; do
  echo "== $cfg"; env $cfg CUGRAPH_B200_SSSP_TRACE=1 timeout 120 $B 24 trav 2 2>gpurun_out/trace.tmp | grep sssp; grep "sssp window" gpurun_out/trace.tmp | tail -1
done 2>&1 | tee gpurun_out/r02_sssp_knobs.log
