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
# the bench command launches the sweep 100 x per step: skip past the first PageRank call, capture one k_sweep + one k_sweep_finish
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:k_sweep" -s 40 -c 2 -f \
  -o gpurun_out/${TAG}_prof_sweep env CUGRAPH_B200_BENCH_TRAVERSAL=0 python bench.py --steps 1 --warmup 1 > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | tail -5
