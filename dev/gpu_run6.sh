#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_nets.py -m gpu -q --no-header -rf -x --timeout 300 > gpurun_out/r2_t6a.log 2>&1; tail -15 gpurun_out/r2_t6a.log
python dev/sweep_env.py "FR_TC_2CTA=0" "FR_TC_2CTA=1" "FR_TC_2CTA=3" "FR_TC_2CTA=5" "FR_TC_2CTA=7" 2>&1 | tee gpurun_out/r2_sweep4.txt
