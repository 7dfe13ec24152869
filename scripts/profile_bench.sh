#!/bin/bash
# ncu evidence for profiles/ in ONE GPU call (single GPU):
#   gpurun --timeout 900 -- 'bash scripts/profile_bench.sh r02'
# 1. launch list of the bench command itself (my kernels only; staging kernels of graph creation included, capped)
# 2. full-set capture of the sweep kernels (one launch each, taken after the warm-up sweeps)
# then:  python scripts/summarize_ncu.py r02     (on the CPU box, reads gpurun_out/)
TAG=${1:-rXX}
mkdir -p gpurun_out
make -s -C oracle
timeout 60 python -c "import torch; torch.zeros(1, device='cuda'); print('cuda ok')"
# a bench value printed under ncu is never a bench value: steps 1, warmup 1, output discarded
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:^k_ -c 600 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 1 --warmup 1 \
  > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:k_spmv_blocked(_x)?$|k_spmv_low" -s 6 -c 2 -f \
  -o gpurun_out/${TAG}_prof_sweep python scripts/quick_bench.py 24 3 > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | tail -5
