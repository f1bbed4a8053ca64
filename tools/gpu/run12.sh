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
    print(t, "value %.0f e2e %.0f ok %.3f launches/frame %.2f"%(d["value"], d["e2e"]["value"], d["config"]["tracked_ok_frac"], d["gpu_launches_per_frame"]))
except Exception as e:
    print(t, "ERR", e); print(open("gpurun_out/r2_ab_%s.err"%t).read()[-600:])
PY
}
run base A=1
run fmfused REBVO_B200_FM_FUSED=1
run base2 A=1
