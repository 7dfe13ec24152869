#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02_sssp_ab.log
: > $out
run() { echo "== $*" >> $out; env "$@" timeout 200 ./cugraph_b200/lib/cbench 24 trav 4 2>&1 | grep -E "^\{|window" | tail -4 >> $out; }
run CUGRAPH_B200_SSSP_SMALL_ROUNDS=0 CUGRAPH_B200_SSSP_TRACE=1
run CUGRAPH_B200_SSSP_SMALL_ROUNDS=1 CUGRAPH_B200_SSSP_TRACE=1
run A=1
cat $out
timeout 600 python -m pytest tests/test_traversal_gpu.py -x -q 2>&1 | tail -3
