#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scalespace_edges.py tests/test_gpu_pipeline.py tests/test_gpu_undistort.py -x -q 2>&1 | tail -12
timeout 600 python bench.py > gpurun_out/bench_und.json 2>gpurun_out/bench_und.err; tail -c 400 gpurun_out/bench_und.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_und.json').read().strip().splitlines()[-1]); ss=d['roofline']['scale_space']
print('value %.0f e2e %.0f cpu %s'%(d['value'], d['e2e']['value'], d['cpu_baseline']['value']))
print(d['parity'])
print({k:(round(v['ms_per_launch']*1e3,1), round(v['frac'],3)) for k,v in ss['all_passes'].items()}, round(ss['whole_scale_space_frac'],3))
PY
