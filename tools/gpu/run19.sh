#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline | tee gpurun_out/bench_rowtma.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); ss=d['roofline']['scale_space']
print('value %.0f e2e %.0f'%(d['value'], d['e2e']['value']), ss['kernel'], round(ss['frac'],3), round(ss['whole_scale_space_frac'],3), {k:round(v['frac'],3) for k,v in ss['all_passes'].items()})"
