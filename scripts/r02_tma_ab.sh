#!/bin/bash
# NOTE: records how profiles/r02_tma_ab.log was produced, on a tree with profiles/r02_tma_stream_attempt.patch applied (not kept).
mkdir -p gpurun_out
out=gpurun_out/r02_tma_ab.log
: > $out
run() { echo "== $*" >> $out; env "$@" CUGRAPH_B200_BUILD_TRACE=1 timeout 300 ./cugraph_b200/lib/cbench 24 sweep >> $out 2>&1; }
run A=0
run CUGRAPH_B200_SWEEP_ALIGN=1
run CUGRAPH_B200_SWEEP_ALIGN=1 CUGRAPH_B200_SWEEP_COST_SLOT=20
grep -E "^==|sweep_ms|step-rows|piece slots|layout: fill" $out
timeout 600 python -m pytest tests/test_pagerank_gpu.py -x -q 2>&1 | tail -5
