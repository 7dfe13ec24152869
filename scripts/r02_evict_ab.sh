#!/bin/bash
# NOTE: the three library variants compared here (B200_ACC_POLICY = 0 / 1 / 2) existed only for this experiment; the kept code is variant 2.
# A/B of the L2 eviction-policy builds of the sweep (profiles/r02_evict_ab3.log); earlier contents of this scratch script produced
# r02_tma_ab / r02_fullchunk_ab / r02_sssp_ab / r02_evict_ab{,2}.log
mkdir -p gpurun_out
out=gpurun_out/r02_evict_ab3.log
: > $out
for rep in 1 2; do
for v in lib_a lib lib_c; do
  echo "== $v (a: plain RED, lib: run-time choice off, c: always evict-last) sweep" >> $out
  timeout 120 ./cugraph_b200/$v/cbench 24 sweep >> $out 2>&1
done
done
for v in lib_a lib_c; do
  echo "== $v pagerank" >> $out
  timeout 120 ./cugraph_b200/$v/cbench 24 pagerank >> $out 2>&1
done
grep -E "^==|sweep_ms|pagerank100|rror" $out | sed 's/"max_rel.*//'
