#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_nets.py -m gpu -q --no-header -rf --timeout 300 > gpurun_out/r2_t12.log 2>&1; tail -5 gpurun_out/r2_t12.log
python dev/sweep_env.py "FR_TC_EPIALT=0" "FR_TC_EPIALT=64" "FR_TC_EPIALT=128" "FR_TC_EPIALT=128 FR_TC_TMASTORE=0" 2>&1 | tee gpurun_out/r2_sweep9.txt
