#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py -x -q -k "undistort" 2>&1 | tail -3
for k in 2 4 8; do
REBVO_B200_UG_IMGS=$k timeout 600 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/bench_und.json 2>gpurun_out/bench_und.err; tail -c 400 gpurun_out/bench_und.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_und.json').read().strip().splitlines()[-1]); ss=d['roofline']['scale_space']
print('IMGS=$k value %.0f e2e %.0f'%(d['value'], d['e2e']['value']), ss['all_passes']['k_undistort_gray']['ms_per_launch']*1e3, round(ss['whole_scale_space_frac'],3))
PY
done
