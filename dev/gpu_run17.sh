#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_nets.py -m gpu -q --no-header -rf --timeout 600 -k "k6 or iresnet or embed" > gpurun_out/r2_t17.log 2>&1; tail -6 gpurun_out/r2_t17.log
python dev/sweep_env.py "FR_X=1" "FR_X=2" 2>&1 | tee gpurun_out/r2_sweep17.txt
ncu --set full --clock-control none --import-source on -k regex:stem_mma -s 2 -c 1 -o gpurun_out/r2_stem_full python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gallery > gpurun_out/ncu_stem.log 2>&1; tail -2 gpurun_out/ncu_stem.log
