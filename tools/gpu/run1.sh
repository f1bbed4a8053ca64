#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo "rc=$?" >> gpurun_out/r2_bench_a.err
REBVO_B200_MIN_XCHG=0 timeout 600 python bench.py --no-cpu-baseline --steps 4 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
REBVO_B200_MIN_CLUSTER=0 timeout 600 python bench.py --no-cpu-baseline --steps 4 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err
tail -3 gpurun_out/r2_pytest1.log
python - <<'PY'
import json
for n in "abc":
    try:
        d=json.loads(open("gpurun_out/r2_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, d["value"], d["e2e"]["value"], d["roofline"]["time_dominant"]["stage_us_per_frame_eager"])
    except Exception as e:
        print(n, "ERR", e)
PY
