#!/bin/bash
# First GPU call of the next round: parity of every experimental path, then an interleaved A/B of their speed.
#   gpurun --timeout 900 -- 'bash scripts/round2_first_call.sh'
# Each step has its own timeout (a cold box once stalled python start-up for minutes, profiles/r01_notes.md).
mkdir -p gpurun_out
make -s -C oracle
timeout 60 python -c "import torch; torch.zeros(1, device='cuda'); print('cuda ok')"          # warms the image
timeout 400 bash scripts/test_variants.sh 2>&1 | tee gpurun_out/variants_parity.log | grep -E "^==|passed|failed|error" 
X=CUGRAPH_B200_HOT_X=1
timeout 400 python scripts/sweep_knobs.py 24 2 - $X $X,CUGRAPH_B200_HOT_NARROW=1 CUGRAPH_B200_LOW_ELL=1 CUGRAPH_B200_LOW_ELL=2 \
  CUGRAPH_B200_LOW_ASYNC=1 $X,CUGRAPH_B200_HOT_NARROW=1,CUGRAPH_B200_LOW_ELL=2 \
  $X,CUGRAPH_B200_HOT_NARROW=1,CUGRAPH_B200_LOW_ELL=2,CUGRAPH_B200_LOW_ASYNC=1 \
  $X,CUGRAPH_B200_HOT_NARROW=1,CUGRAPH_B200_HOT_MIN_DEGREE=8 $X,CUGRAPH_B200_HOT_NARROW=1,CUGRAPH_B200_HOT_MIN_DEGREE=1 \
  $X,CUGRAPH_B200_HOT_CLAIM=8,CUGRAPH_B200_HOT_UNIT_SLOTS=4096 $X,CUGRAPH_B200_HOT_CLAIM=2,CUGRAPH_B200_HOT_UNIT_SLOTS=16384 \
  2>&1 | tee gpurun_out/variants_speed.log | tail -12
CUGRAPH_B200_BUILD_TRACE=1 timeout 120 python scripts/quick_e2e.py 24 2 2>&1 | tee gpurun_out/e2e_trace.log | tail -24
