#!/bin/bash
mkdir -p gpurun_out
B=./cugraph_b200/lib/cbench
for r in 1 2; do
echo "== bank order on"; timeout 120 $B 24 sweep | sed 's/"max_rel.*//'
echo "== bank order off"; CUGRAPH_B200_SWEEP_BANK_ORDER=0 timeout 120 $B 24 sweep | sed 's/"max_rel.*//'
done
CUGRAPH_B200_BUILD_TRACE=1 timeout 200 python scripts/quick_e2e.py 24 2 2>&1 | tail -45
