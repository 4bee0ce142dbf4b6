#!/bin/bash
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gallery"
export FR_GRAPHS=0
ncu --set full --clock-control none --import-source on -k regex:conv3_halo -s 5 -c 1 -o gpurun_out/r2_c3_h0out_full $CMD > gpurun_out/ncu_c3b.log 2>&1; tail -1 gpurun_out/ncu_c3b.log
ncu --set full --clock-control none --import-source on -k regex:sep_gemm -s 7 -c 1 -o gpurun_out/r2_sep_s31_full $CMD > gpurun_out/ncu_sepb.log 2>&1; tail -1 gpurun_out/ncu_sepb.log
