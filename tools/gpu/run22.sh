#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
for g in 4 5 6 3; do
REBVO_B200_MIN_G=$g timeout 300 python bench.py --no-cpu-baseline --steps 6 > /tmp/b.json
python - <<PY
import json
d=json.loads(open('/tmp/b.json').read().strip().splitlines()[-1])
print('G=$g value %.0f e2e %.0f min_us %.1f'%(d['value'], d['e2e']['value'], d['roofline']['stage_us_per_frame_eager']['minimizer']))
PY
done
