#!/bin/bash
for g in 0 1 0 1; do
  FR_GRAPHS=$g python bench.py --steps 20 --warmup 16 --no-cpu-baseline --no-gallery 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('FR_GRAPHS=$g', 'step %.3f'%d['ms_per_step'], 'value %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], 'instr %.3f'%d['detail']['ms_per_step_instrumented_pass'], d['clocks']['sm_mhz'])"
done
