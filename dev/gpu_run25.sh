#!/bin/bash
for k in 1 2 3 4 5 6; do
  python bench.py --no-cpu-baseline --no-gallery 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('run $k', 'step %.3f'%d['ms_per_step'], 'value %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], 'instr %.3f'%d['detail']['ms_per_step_instrumented_pass'], d['clocks'])"
done
