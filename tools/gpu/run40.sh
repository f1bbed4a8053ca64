#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2e_bench_default.json 2> gpurun_out/r2e_bench_default.err; tail -c 300 gpurun_out/r2e_bench_default.err
timeout 900 python bench.py --config 5 > gpurun_out/r2e_bench_config5.json 2>/dev/null
timeout 900 python bench.py --config 4 --seqs 8 > gpurun_out/r2e_bench_config4.json 2>/dev/null
timeout 900 python bench.py --config 3 > gpurun_out/r2e_bench_config3.json 2>/dev/null
for f in default config3 config5 config4; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2e_bench_$f.json').read().strip().splitlines()[-1])
    print('$f', round(d['value'],1), round(d.get('e2e',{}).get('value',0),1), (d.get('cpu_baseline') or {}).get('value'), (d.get('parity') or {}).get('ate_m'))
except Exception as e: print('$f failed', e)
PY
done
