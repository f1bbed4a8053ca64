#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
for cfg in "1 0" "0 0" "1 16" "1 32"; do
set -- $cfg
REBVO_B200_PRIO=$1 REBVO_B200_SS_SUB=$2 timeout 600 python bench.py --no-cpu-baseline --steps 6 > gpurun_out/bench_ss.json 2>gpurun_out/bench_ss.err; tail -c 300 gpurun_out/bench_ss.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_ss.json').read().strip().splitlines()[-1])
print('PRIO=$1 SS_SUB=$2 value %.0f e2e %.0f mirror %.0f'%(d['value'], d['e2e']['value'], d['e2e_with_mirror']['keyline_168B']['value']))
PY
done
