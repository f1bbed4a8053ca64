#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_minimizer_cluster -s 4 -c 1 -f -o gpurun_out/prof_mc python tools/run_few.py 8 > gpurun_out/ncu_mc.log 2>&1
tail -5 gpurun_out/ncu_mc.log
