#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_shim.py tests/test_gpu_clone.py -q -s > gpurun_out/r2_pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest4.log
grep -v "^$" gpurun_out/r2_pytest4.log | tail -14
