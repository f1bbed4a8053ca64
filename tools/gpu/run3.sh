#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
timeout 120 python tools/cluster_stamps.py 2>&1 | grep -v "^   round" > gpurun_out/r2_stamps7.log
cat gpurun_out/r2_stamps7.log
