#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
for q in 1 0 1 0 1 0; do
REBVO_B200_POST_IN_FM=$q timeout 600 python bench.py --no-cpu-baseline --steps 8 > gpurun_out/bench_q.json 2>gpurun_out/bench_q.err; tail -c 300 gpurun_out/bench_q.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_q.json').read().strip().splitlines()[-1])
print('POST_IN_FM=$q value %.0f e2e %.0f'%(d['value'], d['e2e']['value']))
PY
done
