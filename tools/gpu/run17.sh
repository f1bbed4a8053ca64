#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py tests/test_gpu_branches.py -x -q 2>&1 | tail -5
for m in 1 0; do
REBVO_B200_ROW_TMA=$m timeout 300 python bench.py --no-cpu-baseline --steps 6 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ROW_TMA=$m value %.0f e2e %.0f'%(d['value'], d['e2e']['value'])); print(json.dumps(d['roofline'].get('scale_space', d['roofline']))[:1500])"
done
