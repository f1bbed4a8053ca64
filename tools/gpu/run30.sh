#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_shim.py tests/test_gpu_undistort.py tests/test_gpu_dropin.py tests/test_gpu_imu_mode.py -x -q 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r2c_bench_2gpu.json 2> gpurun_out/r2c_bench_2gpu.err; tail -c 300 gpurun_out/r2c_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c_bench_2gpu.json').read().strip().splitlines()[-1])
print('2gpu value %.0f e2e %.0f n_gpus %d'%(d['value'], d['e2e']['value'], d['n_gpus']))
PY
