#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py -m gpu -q --no-header -rf --timeout 600 > gpurun_out/r2_t21.log 2>&1; tail -6 gpurun_out/r2_t21.log
python dev/sweep_env.py "FR_X=1" "FR_X=2" 2>&1 | tee gpurun_out/r2_sweep21.txt
SWEEP_STEPS=20 python dev/sweep_env.py "FR_X=1" "FR_X=2" 2>&1 | tee -a gpurun_out/r2_sweep21.txt
