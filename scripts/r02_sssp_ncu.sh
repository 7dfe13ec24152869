#!/bin/bash
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_trav_launches.csv ./cugraph_b200/lib/cbench 24 trav 1 > gpurun_out/r02_trav_ncu.log 2>&1
tail -3 gpurun_out/r02_trav_ncu.log; wc -l gpurun_out/r02_trav_launches.csv
