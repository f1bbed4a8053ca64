#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
timeout 300 python tools/cluster_stamps.py > gpurun_out/r2_stamps.log 2>&1
timeout 300 python tools/trace_run.py > gpurun_out/r2_trace.log 2>&1
cat gpurun_out/r2_stamps.log; tail -5 gpurun_out/r2_trace.log
