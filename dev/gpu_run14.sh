#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py -m gpu -q --no-header -rf --timeout 300 > gpurun_out/r2_t14.log 2>&1; tail -8 gpurun_out/r2_t14.log
python dev/sweep_env.py "FR_SCRFD_FRONT_FUSED=1" "FR_SCRFD_FRONT_FUSED=1" 2>&1 | tee gpurun_out/r2_sweep14.txt
