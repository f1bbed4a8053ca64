#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_imu_mode.py -q -s 2>&1 | grep -v "^$" | tail -5 | cut -c1-200
for c in 3 5; do
  timeout 900 python bench.py --config $c --steps 4 > gpurun_out/r2_c$c.json 2> gpurun_out/r2_c$c.err
  python - $c <<'PY'
import json,sys
c=sys.argv[1]
try:
    d=json.loads(open("gpurun_out/r2_c%s.json"%c).read().strip().splitlines()[-1])
    print("config", c, "value %.0f e2e %.0f"%(d["value"], d["e2e"]["value"]), "cpu", d["cpu_baseline"]["value"], "parity", {k:d["parity"][k] for k in ("frames","ate_m","max_pos_err_m","kn_equal","matches_equal")} if d["parity"] else None)
except Exception as e:
    print(c, "ERR", e); print(open("gpurun_out/r2_c%s.err"%c).read()[-800:])
PY
done
timeout 600 python bench.py --config 3 --impl reference --steps 2 --warmup 1 | cut -c1-400
