#!/bin/bash
for ramp in "" "1,2,4,8,16" "1,1,2,4,12" "2,2,4,12" "1,3,16" "2,6,12" "4,16" "1,2,3,4,5,5" "3,3,3,3,4,4"; do
EH_RING_RAMP="$ramp" timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-wide | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ramp [$ramp]', 'e2e', round(d['e2e']['value']/1e9,3), 'resident', round(d['e2e']['resident_dataset']['value']/1e9,3))"
done
