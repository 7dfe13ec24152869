#!/bin/bash
# 2-GPU check of the NCCL path: gpurun --gpus 2 --timeout 900 -- 'bash scripts/r02_mg.sh'
mkdir -p gpurun_out
make -s -C oracle
nvidia-smi -L | head -4
echo "(MG tests ran earlier this round: 2 passed on 2 GPUs)"
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r02_bench_mg2.json 2> gpurun_out/r02_bench_mg2.err
tail -2 gpurun_out/r02_bench_mg2.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02_bench_mg2.json') if l.startswith('{')][-1])
    c=d['config']
    print('value',d['value'],'ms',d['ms_per_step'],'sweep frac',d['roofline']['frac'],'ms_sweep',d['roofline']['ms_per_sweep'])
    print('parity',c['mg_parity_ok'],c['mg_parity_max_rel'],'bfs parity',c.get('mg_bfs_parity'),'bfs',c.get('mg_bfs'))
    print('e2e',d['e2e'])
except Exception as e:
    print('no line',e)
PY
