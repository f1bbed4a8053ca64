#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
for ns in 3 5 6 8; do
REBVO_B200_ROW_NS=$ns timeout 300 python bench.py --no-cpu-baseline --steps 4 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); ss=d['roofline']['scale_space']['all_passes']
print('NS=$ns value %.0f'%d['value'], {k:round(v['ms_per_launch']*1e3,1) for k,v in ss.items()})"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rowscan_tma_avg --launch-skip 4 -c 2 -o gpurun_out/r2_prof_rowavg -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/prof_row.log 2>&1
tail -2 gpurun_out/prof_row.log | cut -c1-200
