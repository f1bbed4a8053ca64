#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_imu_mode.py -q -s > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest6.log
grep -v "^$" gpurun_out/r2_pytest6.log | tail -30 | cut -c1-250
