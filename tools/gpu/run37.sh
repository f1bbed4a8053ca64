#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
for q in 64 128 96; do
timeout 600 python bench.py --no-cpu-baseline --steps 8 --batch $q > gpurun_out/bench_q.json 2>gpurun_out/bench_q.err; tail -c 300 gpurun_out/bench_q.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_q.json').read().strip().splitlines()[-1])
print('B=$q value %.0f e2e %.0f mirror %.0f'%(d['value'], d['e2e']['value'], d['e2e_with_mirror']['keyline_168B']['value']))
PY
done
