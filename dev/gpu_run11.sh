#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_nets.py -m gpu -q --no-header -rf --timeout 300 > gpurun_out/r2_t11.log 2>&1; tail -5 gpurun_out/r2_t11.log
python dev/sweep_env.py "FR_X=0" "FR_X=1" 2>&1 | tee gpurun_out/r2_sweep8.txt
