#!/bin/bash
# Round 2, first GPU call: hardware numbers of what round 1 left unmeasured.
mkdir -p gpurun_out
make -s -C oracle
B=./cugraph_b200/lib/cbench
X=CUGRAPH_B200_HOT_X=1; N=CUGRAPH_B200_HOT_NARROW=1; K=CUGRAPH_B200_HOT_BANK_ORDER=1
{
echo "== defaults"; timeout 60 $B 24 sweep
echo "== X N K LOW_ELL=2"; env $X $N $K CUGRAPH_B200_LOW_ELL=2 timeout 60 $B 24 sweep
echo "== X N K MD1"; env $X $N $K CUGRAPH_B200_HOT_MIN_DEGREE=1 timeout 60 $B 24 sweep
echo "== X N K MD8 LOW_ELL=2"; env $X $N $K CUGRAPH_B200_HOT_MIN_DEGREE=8 CUGRAPH_B200_LOW_ELL=2 timeout 60 $B 24 sweep
echo "== traversal defaults"; timeout 90 $B 24 trav 8
echo "== traversal BFS alpha 40"; env CUGRAPH_B200_BFS_ALPHA=40 timeout 90 $B 24 trav 8
echo "== traversal BFS alpha 120, SSSP fixed width"; env CUGRAPH_B200_BFS_ALPHA=120 CUGRAPH_B200_SSSP_ADAPTIVE=0 timeout 90 $B 24 trav 8
} 2>&1 | tee gpurun_out/r02_call1_cbench.log
# ncu full of the best all-rows variant and of the ELL-hot low kernel
env $X $N $K CUGRAPH_B200_HOT_MIN_DEGREE=1 timeout 300 ncu --set full --clock-control none --import-source on \
  -k regex:k_spmv_blocked_x -s 4 -c 1 -f -o gpurun_out/r02_ncu_x_md1 $B 24 sweep > gpurun_out/r02_ncu_x_md1.log 2>&1
env $X $N $K CUGRAPH_B200_LOW_ELL=2 timeout 300 ncu --set full --clock-control none --import-source on \
  -k "regex:k_spmv_blocked_x|k_spmv_low_ell_hot" -s 8 -c 2 -f -o gpurun_out/r02_ncu_x_ell2 $B 24 sweep > gpurun_out/r02_ncu_x_ell2.log 2>&1
timeout 120 python -c "import torch; torch.zeros(1, device='cuda'); print('cuda ok')"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tee gpurun_out/r02_pytest_gpu0.log | tail -5
CUGRAPH_B200_SSSP_TRACE=1 CUGRAPH_B200_BFS_TRACE=1 timeout 200 python scripts/quick_trav.py 24 2 2>&1 | tee gpurun_out/r02_trav_trace.log | tail -40
ls -la gpurun_out | tail
