#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --steps 6 > gpurun_out/r2_ab_$tag.json 2> gpurun_out/r2_ab_$tag.err
  python - "$tag" <<'PY'
import json,sys
t=sys.argv[1]
try:
    d=json.loads(open("gpurun_out/r2_ab_%s.json"%t).read().strip().splitlines()[-1])
    print(t, "value %.0f e2e %.0f min %.1f ok %.3f"%(d["value"], d["e2e"]["value"], d["roofline"]["time_dominant"]["stage_us_per_frame_eager"]["minimizer"], d["config"]["tracked_ok_frac"]))
except Exception as e:
    print(t, "ERR", e); print(open("gpurun_out/r2_ab_%s.err"%t).read()[-600:])
PY
}
run g4 A=1
run g3 REBVO_B200_MIN_G=3
run g5 REBVO_B200_MIN_G=5
run g6 REBVO_B200_MIN_G=6
run g4b A=1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
