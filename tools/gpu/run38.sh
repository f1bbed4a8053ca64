#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py -x -q -k "undistort" 2>&1 | tail -5
for q in 1 0; do
REBVO_B200_UG_TILED=$q timeout 600 python bench.py --no-cpu-baseline --steps 4 > gpurun_out/bench_q.json 2>gpurun_out/bench_q.err; tail -c 300 gpurun_out/bench_q.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_q.json').read().strip().splitlines()[-1]); ss=d['roofline']['scale_space']
print('TILED=$q value %.0f e2e %.0f'%(d['value'], d['e2e']['value']), ss['all_passes']['k_undistort_gray'], round(ss['whole_scale_space_frac'],3))
PY
done
