#!/bin/bash
# multi-GPU check: NCCL tests + the N-GPU bench line     gpurun --gpus N -- 'bash scripts/r02_mg.sh N'
N=${1:-2}
mkdir -p gpurun_out
make -s -C oracle
timeout 600 python -m pytest tests/test_mg_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/r02_bench_mg$N.json 2> gpurun_out/r02_bench_mg$N.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02_bench_mg$N.json') if l.startswith('{')][-1])
    print('value',d['value'],'ms/step',d['ms_per_step'],'roofline',d['roofline']['frac'],d['roofline']['ms_per_sweep'])
    print({k:v for k,v in d['config'].items() if k.startswith(('mg_','mass','workload'))})
    print('e2e',d['e2e'])
except Exception as e:
    print('no bench line',e)
PY
tail -5 gpurun_out/r02_bench_mg$N.err
