#!/bin/bash
for ns in 0 200 1000 0 200 1000; do
  FR_POLL_NS=$ns python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-gallery 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('POLL_NS=$ns', 'step %.3f'%d['ms_per_step'], 'value %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], d['clocks']['sm_mhz'], 'trunk %.3f scrfd %.3f'%(d['detail']['stage_ms_per_step']['trunk'], d['detail']['stage_ms_per_step']['scrfd']))"
done
