#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_nets.py tests/test_gpu_pipeline.py -m gpu -q --no-header -rf --timeout 600 > gpurun_out/r2_t8.log 2>&1; tail -12 gpurun_out/r2_t8.log
python dev/sweep_env.py "FR_TC_TMASTORE=0" "FR_TC_TMASTORE=1" "FR_TC_TMASTORE=1 FR_TC_ASTAGES=3" 2>&1 | tee gpurun_out/r2_sweep6.txt
