#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
for k in 1 2; do timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -2; done
