#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_full.log
tail -4 gpurun_out/r2_pytest_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -c 300 gpurun_out/r2_bench_final.err
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_final_ref.json 2>> gpurun_out/r2_bench_final.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_final.json").read().strip().splitlines()[-1])
r=json.loads(open("gpurun_out/r2_bench_final_ref.json").read().strip().splitlines()[-1])
print("ours", d["value"], d["e2e"]["value"], "ref", r["value"], "ratio e2e", d["e2e"]["value"]/r["value"], d["clocks"])
print(d["parity"]["ate_m"], d["roofline"]["frac"], d["roofline"]["scale_space"]["whole_scale_space_frac"])
PY
