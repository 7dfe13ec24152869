#!/bin/bash
# A/B of the sweep variants and the traversal schedule knobs with the Python-free probe: a dozen configurations in well
# under a minute of GPU time.     gpurun --timeout 240 -- 'bash scripts/cbench_ab.sh 24 2>&1 | tee gpurun_out/cbench_ab.log'
S=${1:-24}
B=./cugraph_b200/lib/cbench
X=CUGRAPH_B200_HOT_X=1; N=CUGRAPH_B200_HOT_NARROW=1; K=CUGRAPH_B200_HOT_BANK_ORDER=1
run() { echo "== ${*:-defaults}"; env "$@" timeout 40 $B $S sweep; }
for round in 1 2; do   # two interleaved rounds: drift shows up between rounds, not between configurations
  run
  run $X
  run $K
  run $X $K
  run $X $N
  run $X $N $K
  run CUGRAPH_B200_LOW_ELL=1
  run CUGRAPH_B200_LOW_ELL=2
  run $X $N $K CUGRAPH_B200_LOW_ELL=2
  run $X $N CUGRAPH_B200_HOT_MIN_DEGREE=8
  run $X $N $K CUGRAPH_B200_HOT_MIN_DEGREE=1
  run CUGRAPH_B200_LOW_ASYNC=1
done
echo "== pagerank defaults"; timeout 40 $B $S pagerank
echo "== traversal defaults"; timeout 60 $B $S trav 8
echo "== traversal BFS alpha 40"; env CUGRAPH_B200_BFS_ALPHA=40 timeout 60 $B $S trav 8
echo "== traversal BFS alpha 120, SSSP fixed width"; env CUGRAPH_B200_BFS_ALPHA=120 CUGRAPH_B200_SSSP_ADAPTIVE=0 timeout 60 $B $S trav 8
