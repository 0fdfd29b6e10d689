#!/bin/bash
for cfg in "A=1" "EH_USE_X2=1" "FLAGS=16" "EH_TC_MIN_BATCH=16384"; do
fl=0; if [ "$cfg" = "FLAGS=16" ]; then fl=16; fi
env $cfg timeout 120 python bench.py --steps 1024 --warmup 32 --no-cpu-baseline --no-wide --no-e2e --flags $fl | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$cfg', 'us/step', round(d['ms_per_step']*1e3,3), d['final_loss'])"
done
