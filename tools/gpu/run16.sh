#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
timeout 120 python tools/cluster_stamps.py 2>&1 | grep -v "CTA 15" | head -14
timeout 300 python bench.py --no-cpu-baseline --steps 6 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.0f e2e %.0f'%(d['value'], d['e2e']['value']))"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -x -q 2>&1 | tail -2
