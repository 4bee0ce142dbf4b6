#!/bin/bash
# final round-2 measurement, part B: ncu passes of the command that part A ran clean (same binary, same args
# except the shorter step count)
# FR_GRAPHS=0: the launches go out one by one (the kernels are the same ones the graph replays)
export FR_GRAPHS=0
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gallery"
$CMD > gpurun_out/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_step.csv $CMD > gpurun_out/r2_ncu1.log 2>&1
echo "pass1 rc=$?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none -s 270 -c 200 --csv --log-file gpurun_out/r2_step_ncu_metrics.csv $CMD > gpurun_out/r2_ncu2.log 2>&1
echo "pass2 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:halo_gemm2 -s 30 -c 2 -o gpurun_out/r2_halo_gemm2_full $CMD > gpurun_out/r2_ncu3.log 2>&1
echo "pass3 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:front_fused -s 2 -c 1 -o gpurun_out/r2_front_fused_full $CMD > gpurun_out/r2_ncu4.log 2>&1
echo "pass4 rc=$?"
ls -la gpurun_out/ | tail -8
