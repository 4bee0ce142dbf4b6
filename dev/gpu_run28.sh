#!/bin/bash
for spec in "FR_X=0" "FR_TC_2CTA=9" "FR_X=0" "FR_TC_2CTA=9" "FR_TC_ASTAGES2=3" "FR_TC_TAILSPLIT=0"; do
  env $spec python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-gallery 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$spec', 'step %.3f'%d['ms_per_step'], 'value %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], d['clocks']['sm_mhz'], 'trunk %.3f scrfd %.3f'%(d['detail']['stage_ms_per_step']['trunk'], d['detail']['stage_ms_per_step']['scrfd']))"
done
