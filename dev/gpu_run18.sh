#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_nets.py tests/test_gpu_onnx.py -m gpu -q --no-header -rf --timeout 600 -k "scrfd or det or onnx" > gpurun_out/r2_t18.log 2>&1; tail -6 gpurun_out/r2_t18.log
python dev/sweep_env.py "FR_SCRFD_TAIL8=0" "FR_SCRFD_TAIL8=1" "FR_SCRFD_TAIL8=0" "FR_SCRFD_TAIL8=1" 2>&1 | tee gpurun_out/r2_tail8_sweep.txt
