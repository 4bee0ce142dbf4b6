#!/bin/bash
# fused SCRFD front (stem -> b0 -> s0.0): parity tests, A/B against the three separate kernels, one ncu capture
set -x
timeout 900 python -m pytest tests/test_gpu_nets.py -m gpu -q --no-header -rf --timeout 300 -k "scrfd" > gpurun_out/r2_t13.log 2>&1; tail -15 gpurun_out/r2_t13.log
python dev/sweep_env.py "FR_SCRFD_FRONT_FUSED=0" "FR_SCRFD_FRONT_FUSED=1" "FR_SCRFD_FRONT_FUSED=0" "FR_SCRFD_FRONT_FUSED=1" 2>&1 | tee gpurun_out/r2_front_fused_sweep.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:front_fused -c 1 -o gpurun_out/front_fused python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gallery > gpurun_out/ncu13.log 2>&1; tail -3 gpurun_out/ncu13.log
