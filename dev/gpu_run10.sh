#!/bin/bash
set -x
python dev/sweep_env.py "FR_TC_MT=0" "FR_TC_MT=2" "FR_TC_RESB=0" "FR_TC_ASTAGES=4 FR_TC_TMASTORE=0" "FR_TC_ASTAGES=1" 2>&1 | tee gpurun_out/r2_sweep7.txt
for sp in 4 9; do
  FR_GALLERY_SPLITS=$sp python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_gal_$sp.json 2>/dev/null
  python -c "
import json; g=json.load(open('gpurun_out/r2_gal_$sp.json'))['gallery_1toN']; print('splits $sp', g['value'], g['ms_per_batch'], g['fp8']['value'], g['fp8']['ms_per_batch'], g['fp8']['agreement_with_bf16_top10'])"
done
timeout 600 python -m pytest tests/test_gpu_gallery.py -m gpu -q --no-header -rf --timeout 600 > gpurun_out/r2_t10.log 2>&1; tail -3 gpurun_out/r2_t10.log
