#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_netpack.py "tests/test_gpu_parity.py::test_stage_parity" -q > gpurun_out/r2_pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest5.log
grep -v "^$" gpurun_out/r2_pytest5.log | tail -40 | cut -c1-220
