#!/bin/bash
# Parity of the experimental code paths (not part of tests/: they are off by default).  Run on a GPU box:
#   bash scripts/test_variants.sh
set -u
T="tests/test_pagerank_gpu.py tests/test_edge_cases_gpu.py"
run() { echo "== $*"; env CUGRAPH_B200_HOT_MIN_EDGES=0 "$@" python -m pytest $T -x -q 2>&1 | tail -2; }
run CUGRAPH_B200_HOT_X=1
run CUGRAPH_B200_HOT_X=1 CUGRAPH_B200_HOT_C1=0 CUGRAPH_B200_HOT_CLAIM=8
run CUGRAPH_B200_HOT_X=1 CUGRAPH_B200_HOT_UNIT_SLOTS=2048
run CUGRAPH_B200_LOW_ELL=1
run CUGRAPH_B200_LOW_ELL=1 CUGRAPH_B200_HOT_MIN_EDGES=1000000000
run CUGRAPH_B200_HOT_BLOCKS=1
run CUGRAPH_B200_LOW_MODE=0
run CUGRAPH_B200_HOT_NARROW=1
run CUGRAPH_B200_HOT_NARROW=1 CUGRAPH_B200_LOW_ELL=1 CUGRAPH_B200_HOT_CLAIM=2
run CUGRAPH_B200_LOW_ASYNC=1
