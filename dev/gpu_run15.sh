#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_cli.py tests/test_capi_host.py -m gpu -q --no-header -rf --timeout 300 > gpurun_out/r2_t15.log 2>&1; tail -12 gpurun_out/r2_t15.log
python dev/sweep_env.py "FR_GRAPHS=0" "FR_GRAPHS=1" 2>&1 | tee gpurun_out/r2_graph_sweep.txt
