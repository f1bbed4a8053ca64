#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py tests/test_gpu_branches.py -x -q 2>&1 | tail -4
for q in 1 0 1 0; do
REBVO_B200_MU_XCHG=$q timeout 600 python bench.py --no-cpu-baseline --steps 8 > gpurun_out/bench_q.json 2>gpurun_out/bench_q.err; tail -c 300 gpurun_out/bench_q.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_q.json').read().strip().splitlines()[-1])
print('MU_XCHG=$q value %.0f e2e %.0f rescale_us %.1f'%(d['value'], d['e2e']['value'], d['roofline']['stage_us_per_frame_eager']['rescale']))
PY
done
