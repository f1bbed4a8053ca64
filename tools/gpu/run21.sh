#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py -x -q 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_mirror.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_mirror.json').read().strip().splitlines()[-1]); ss=d['roofline']['scale_space']
print('value %.0f e2e %.0f'%(d['value'], d['e2e']['value']))
print(json.dumps(d['e2e_with_mirror']))
PY
