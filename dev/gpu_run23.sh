#!/bin/bash
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gallery"
export FR_GRAPHS=0
ncu --set full --clock-control none --import-source on -k regex:conv3_halo -s 6 -c 1 -o gpurun_out/r2_c3_full $CMD > gpurun_out/ncu_c3.log 2>&1; tail -1 gpurun_out/ncu_c3.log
ncu --set full --clock-control none --import-source on -k regex:sep_gemm -s 0 -c 1 -o gpurun_out/r2_sep_full $CMD > gpurun_out/ncu_sep.log 2>&1; tail -1 gpurun_out/ncu_sep.log
