#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py tests/test_gpu_branches.py -x -q 2>&1 | tail -3
for e in 1 0 1 0; do
REBVO_B200_MIN_EARLY=$e timeout 300 python bench.py --no-cpu-baseline --steps 6 > /tmp/b.json
python - <<PY
import json
d=json.loads(open('/tmp/b.json').read().strip().splitlines()[-1])
print('EARLY=$e value %.0f e2e %.0f'%(d['value'], d['e2e']['value']))
PY
done
