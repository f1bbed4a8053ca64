#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
REBVO_B200_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 2600 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline > gpurun_out/r2b_launches.log 2>&1
tail -2 gpurun_out/r2b_launches.log | cut -c1-200
wc -l gpurun_out/r2b_launches.csv
