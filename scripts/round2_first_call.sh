#!/bin/bash
# First GPU call of the next round: everything that was written without a GPU gets its measurement in one call.
#   gpurun --timeout 1500 -- 'bash scripts/round2_first_call.sh'
# Each step has its own timeout (a cold box once stalled python start-up for minutes, profiles/r01_notes.md).
mkdir -p gpurun_out
make -s -C oracle
timeout 120 python -c "import torch; torch.zeros(1, device='cuda'); print('cuda ok')"          # warms the image
# 0. the cheap part first: Python-free A/B of every sweep variant and the traversal knobs (seconds each)
timeout 400 bash scripts/cbench_ab.sh 24 2>&1 | tee gpurun_out/cbench_ab.log | grep -E '^==|sweep_ms|pagerank100|mean_ms'
# 1. parity: the whole GPU suite (includes the RMAT-24 certificates and the reference's C test programs)
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tee gpurun_out/pytest_gpu.log | tail -5
# 2. the bench line; its `side` block carries BFS / SSSP TEPS, the schedule A/B and parity + timing of every experimental
#    sweep configuration (scripts/bench_side.py), so this one run is the A/B the variants lack
timeout 600 python bench.py > gpurun_out/bench_r02_first.json 2> gpurun_out/bench_r02_first.err; tail -c 600 gpurun_out/bench_r02_first.json
# 3. parity of the experimental paths on the test graphs (off-by-default switches)
timeout 400 bash scripts/test_variants.sh 2>&1 | tee gpurun_out/variants_parity.log | grep -E "^==|passed|failed|error"
# 4. staging / PageRank phase breakdown of the e2e path
CUGRAPH_B200_BUILD_TRACE=1 timeout 120 python scripts/quick_e2e.py 24 2 2>&1 | tee gpurun_out/e2e_trace.log | tail -30
# 5. SSSP / BFS traces of one source each (rounds, windows, relaxations; level directions)
CUGRAPH_B200_SSSP_TRACE=1 CUGRAPH_B200_BFS_TRACE=1 timeout 200 python scripts/quick_trav.py 24 2 2>&1 | tee gpurun_out/trav_trace.log | tail -20
