#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2f_bench_default.json 2> gpurun_out/r2f_bench_default.err; tail -c 300 gpurun_out/r2f_bench_default.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2f_bench_default.json').read().strip().splitlines()[-1])
print('default', round(d['value'],1), round(d['e2e']['value'],1), d['cpu_baseline']['value'], d['parity']['ate_m'], d['parity']['kn_equal'], d['parity']['matches_equal'])
PY
