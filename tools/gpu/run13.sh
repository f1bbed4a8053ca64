#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
timeout 300 python tools/tma_probe.py 2>&1 | tail -2
timeout 900 python -m pytest "tests/test_gpu_parity.py" tests/test_gpu_pipeline.py -x -q 2>&1 | tail -3
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --steps 6 > gpurun_out/r2_ab_$tag.json 2> gpurun_out/r2_ab_$tag.err
  python - "$tag" <<'PY'
import json,sys
t=sys.argv[1]
try:
    d=json.loads(open("gpurun_out/r2_ab_%s.json"%t).read().strip().splitlines()[-1])
    p=d["roofline"]["scale_space"]["all_passes"]
    print(t, "value %.0f e2e %.0f"%(d["value"], d["e2e"]["value"]), {k:(round(v["ms_per_launch"]*1e3,1), round(v["gbs"]/6574.1,3)) for k,v in p.items()}, "whole", round(d["roofline"]["scale_space"]["whole_scale_space_frac"],3))
except Exception as e:
    print(t, "ERR", e); print(open("gpurun_out/r2_ab_%s.err"%t).read()[-600:])
PY
}
run tma A=1
