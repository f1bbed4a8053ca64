#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
for m in 0 48 416 28 216 116 132; do
REBVO_B200_COLSCAN=$m timeout 300 python bench.py --no-cpu-baseline --steps 3 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); ss=d['roofline']['scale_space']['all_passes']
print('COL=$m value %.0f'%d['value'], {k:round(v['ms_per_launch']*1e3,1) for k,v in ss.items()})"
done
REBVO_B200_COLSCAN=216 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
