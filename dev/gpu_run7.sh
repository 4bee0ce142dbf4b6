#!/bin/bash
set -x
FR_TC_2CTA=9 timeout 600 python -m pytest tests/test_gpu_nets.py -m gpu -q --no-header -rf --timeout 300 > gpurun_out/r2_t7a.log 2>&1; tail -8 gpurun_out/r2_t7a.log
FR_TC_2CTA=2 FR_TC_TAILSPLIT=0 timeout 600 python -m pytest tests/test_gpu_nets.py -m gpu -q --no-header -rf --timeout 300 -k "plain or two_cta" > gpurun_out/r2_t7b.log 2>&1; grep -E "passed|failed|FAILED" gpurun_out/r2_t7b.log | tail -15
FR_TC_2CTA=2 timeout 600 python -m pytest tests/test_gpu_nets.py -m gpu -q --no-header -rf --timeout 300 -k "plain or two_cta" > gpurun_out/r2_t7c.log 2>&1; grep -E "passed|failed|FAILED" gpurun_out/r2_t7c.log | tail -15
python dev/sweep_env.py "FR_TC_2CTA=1" "FR_TC_2CTA=9" 2>&1 | tee gpurun_out/r2_sweep5.txt
