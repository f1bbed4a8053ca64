#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
timeout 120 python tools/cluster_stamps.py _a 2>&1 | grep -v "^   round  [1-24-7]" > gpurun_out/r2_stamps5.log
cat gpurun_out/r2_stamps5.log
