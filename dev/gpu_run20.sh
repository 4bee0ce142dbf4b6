#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --no-header -rf --timeout 600 > gpurun_out/r2_t20.log 2>&1; tail -15 gpurun_out/r2_t20.log
python dev/sweep_env.py "FR_GRAPHS=0" "FR_GRAPHS=1" 2>&1 | tee gpurun_out/r2_sweep20.txt
SWEEP_STEPS=20 python dev/sweep_env.py "FR_GRAPHS=0" "FR_GRAPHS=1" 2>&1 | tee -a gpurun_out/r2_sweep20.txt
