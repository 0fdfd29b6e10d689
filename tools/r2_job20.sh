#!/bin/bash
run() { # name, env...
  name=$1; shift
  env "$@" timeout 120 python bench.py --steps 512 --warmup 32 --no-cpu-baseline --no-wide --no-e2e > gpurun_out/r2_j20_$name.json 2>> gpurun_out/r2_j20.err
  python -c "
import json
d=json.load(open('gpurun_out/r2_j20_$name.json')); print('$name', 'us/step', round(d['ms_per_step']*1e3,3), d['final_loss'])"
}
run tc EH_TC_MIN_BATCH=16384
run ffma A=1
run tc20 EH_TC_MIN_BATCH=16384
EH_TC_MIN_BATCH=16384 timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-wide --no-e2e | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('tc 20 steps', 'us/step', round(d['ms_per_step']*1e3,3))"
timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-wide --no-e2e | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ffma 20 steps', 'us/step', round(d['ms_per_step']*1e3,3))"
