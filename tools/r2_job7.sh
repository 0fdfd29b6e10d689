#!/bin/bash
for k in 20 40 100 400; do python bench.py --steps $k --warmup 5 --no-cpu-baseline --no-wide --no-e2e | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('coop   steps',d['steps'],'ms_per_step_us',d['ms_per_step']*1e3,'total_us',d['ms_per_step']*1e3*d['steps'])"; done
for k in 20 100; do EH_NO_COOP=1 python bench.py --steps $k --warmup 5 --no-cpu-baseline --no-wide --no-e2e | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('nocoop steps',d['steps'],'ms_per_step_us',d['ms_per_step']*1e3,'total_us',d['ms_per_step']*1e3*d['steps'])"; done
