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
    print(t, "value %.0f e2e %.0f min %.1f"%(d["value"], d["e2e"]["value"], d["roofline"]["time_dominant"]["stage_us_per_frame_eager"]["minimizer"]))
except Exception as e:
    print(t, "ERR", e)
PY
}
run t384 A=1
run t384_noov REBVO_B200_OVERLAP=0
run t384_big REBVO_B200_MIN_KPC=3000
run t512 REBVO_B200_LIB=$GRAFT_REPO_ROOT/rebvo_b200/alt/t512/librebvo_b200.so
run t512_big REBVO_B200_LIB=$GRAFT_REPO_ROOT/rebvo_b200/alt/t512/librebvo_b200.so REBVO_B200_MIN_KPC=3000
run old REBVO_B200_MIN_CLUSTER=0
