#!/bin/bash
# one full-set ncu capture of k_sweep (and the finish kernel) from the Python-free probe: cbench <scale> sweep
TAG=${1:-r02b}
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k "regex:k_sweep" -s 6 -c 2 -f \
  -o gpurun_out/${TAG}_prof_sweep ./cugraph_b200/lib/cbench 24 sweep > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/${TAG}_prof_sweep.ncu-rep
tail -3 gpurun_out/${TAG}_ncu_full.log
