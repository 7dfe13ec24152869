#!/bin/bash
mkdir -p gpurun_out
make -s -C oracle
B=./cugraph_b200/lib/cbench
{ echo "== sweep"; timeout 120 $B 24 sweep; echo "== trav"; timeout 120 $B 24 trav 8; echo "== trav alpha 40"; CUGRAPH_B200_BFS_ALPHA=40 timeout 120 $B 24 trav 8; } 2>&1 | tee gpurun_out/r02_call3_cbench.log
timeout 120 python -c "import torch; torch.zeros(1, device='cuda'); print('cuda ok')"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tee gpurun_out/r02_pytest_gpu1.log | tail -8
