#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest3.log
tail -25 gpurun_out/r2_pytest3.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -4 gpurun_out/r2_smoke.log
timeout 600 python bench.py > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; tail -c 600 gpurun_out/r2_bench_d.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_d.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["parity"], d["cpu_baseline"])
print({k:(v if not isinstance(v,dict) else '...') for k,v in d["roofline"].items()})
print(d["roofline"]["stages"])
PY
