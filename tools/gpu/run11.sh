#!/bin/bash
cd $GRAFT_REPO_ROOT && mkdir -p gpurun_out
REBVO_B200_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 2600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
tail -2 gpurun_out/r2_ncu_bench.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_minimizer_cluster -s 6 -c 1 -f -o gpurun_out/r2_prof_min python tools/run_few.py 12 > gpurun_out/r2_ncu_min.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"k_rowscan_ring|k_blur_dog|k_colscan|k_rgb2gray" -c 8 -f -o gpurun_out/r2_prof_dog python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu_dog.log 2>&1
ls -la gpurun_out/*.ncu-rep
