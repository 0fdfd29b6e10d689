#!/bin/bash
run() { # name, env...
  name=$1; shift
  env "$@" timeout 120 python bench.py --steps 512 --warmup 32 --no-cpu-baseline --no-wide --no-e2e > gpurun_out/r2_j22_$name.json 2>> gpurun_out/r2_j22.err
  python -c "
import json
d=json.load(open('gpurun_out/r2_j22_$name.json')); print('$name', 'us/step', round(d['ms_per_step']*1e3,3), d['final_loss'])"
  env "$@" EH_EPOCH_DEBUG=gpurun_out/r2_j22_$name.bin EH_PROF_LOG2N=24 timeout 120 python tools/epoch_prof_driver.py 0 32 > /dev/null 2>&1
  python tools/epoch_phase_dump.py gpurun_out/r2_j22_$name.bin 2>&1 | sed -n '11,19p'
}
run ffma A=1
run tc EH_TC_MIN_BATCH=16384
timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-wide --no-e2e | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ffma 20 steps', 'us/step', round(d['ms_per_step']*1e3,3))"
