#!/bin/bash
mkdir -p gpurun_out
make -s -C oracle
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_n1.json') if l.startswith('{')][-1])
c=d['config']
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'sweep_ms',d['roofline']['ms_per_sweep'])
print('e2e',d['e2e'])
print({k:c[k] for k in c if k.startswith(('bfs','sssp','trav','net'))})
print('cpu',d['cpu_baseline'])
PY
tail -3 gpurun_out/r02_bench_n1.err
