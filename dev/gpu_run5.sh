#!/bin/bash
# developer GPU session: 2-CTA kernel tests, gallery / dist tests, A/B sweep, bench with the 1:N block
set -x
timeout 600 python -m pytest tests/test_gpu_nets.py -m gpu -q --no-header -rf -x --timeout 300 > gpurun_out/r2_t5a.log 2>&1; tail -15 gpurun_out/r2_t5a.log
timeout 900 python -m pytest tests/test_gpu_gallery.py tests/test_gpu_dist.py -m gpu -q --no-header -rf --timeout 600 > gpurun_out/r2_t5b.log 2>&1; tail -25 gpurun_out/r2_t5b.log
python dev/sweep_env.py "FR_TC_2CTA=0" "FR_TC_2CTA=1" "FR_TC_2CTA=1 FR_TC_ASTAGES2=3" "FR_TC_2CTA=1 FR_TC_ASTAGES2=4" 2>&1 | tee gpurun_out/r2_sweep3.txt
FR_TC_2CTA=0 python bench.py --no-cpu-baseline > gpurun_out/r2_b5.json 2> gpurun_out/r2_b5.err
python -c "
import json; d=json.load(open('gpurun_out/r2_b5.json')); print(json.dumps(d['gallery_1toN'], indent=1)); print(d['value'], d['e2e']['value'], d['detail']['stage_ms_per_step'])"; tail -3 gpurun_out/r2_b5.err
