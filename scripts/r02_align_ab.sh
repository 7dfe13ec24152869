#!/bin/bash
# NOTE: records how profiles/r02_align_ab.log was produced; the window policy and its knobs were removed from the library afterwards
# (profiles/r02_window_policy.patch), so this script only makes sense on a tree with that patch applied.
# A/B of the row-aligned window policy on RMAT-24 (cbench: device RMAT, sweep parity + live sweep time)
mkdir -p gpurun_out
out=gpurun_out/r02_align_ab.log
: > $out
run() { echo "== $*" >> $out; env "$@" CUGRAPH_B200_BUILD_TRACE=1 timeout 300 ./cugraph_b200/lib/cbench 24 sweep >> $out 2>&1; }
run CUGRAPH_B200_SWEEP_ALIGN=0
run CUGRAPH_B200_SWEEP_ALIGN=1
run CUGRAPH_B200_SWEEP_COST_SLOT=5
run CUGRAPH_B200_SWEEP_COST_SLOT=20
run CUGRAPH_B200_SWEEP_COST_SCAT=220
run CUGRAPH_B200_SWEEP_COST_SCAT=120
grep -E "^==|sweep_ms|step-rows|piece slots|fill|window runs|layout: pieces" $out
