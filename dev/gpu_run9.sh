#!/bin/bash
# ncu --set full of the 64-channel halo kernel (what bounds it?) after a clean plain run of the same command
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gallery"
$CMD > gpurun_out/r2_plain9.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:^halo_gemm_kernel$ -s 36 -c 4 -o gpurun_out/r2_halo64_full $CMD > gpurun_out/r2_ncu9.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/r2_ncu9.log
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --no-header -rf --timeout 600 -k edge > gpurun_out/r2_t9.log 2>&1; tail -3 gpurun_out/r2_t9.log
