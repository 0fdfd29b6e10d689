#!/bin/bash
export EH_TC_MIN_BATCH=16384
run() { # name, env...
  name=$1; shift
  env "$@" timeout 120 python bench.py --steps 512 --warmup 32 --no-cpu-baseline --no-wide --no-e2e > gpurun_out/r2_j15_$name.json 2>> gpurun_out/r2_j15.err
  python -c "
import json
d=json.load(open('gpurun_out/r2_j15_$name.json')); print('$name', 'us/step', round(d['ms_per_step']*1e3,3), d['final_loss'])"
  env "$@" EH_EPOCH_DEBUG=gpurun_out/r2_j15_$name.bin EH_PROF_LOG2N=24 timeout 120 python tools/epoch_prof_driver.py 0 32 > /dev/null 2>&1
  python tools/epoch_phase_dump.py gpurun_out/r2_j15_$name.bin 2>&1 | sed -n '12,14p;19p'
}
run s0 EH_TC_STAGGER_NS=0
run s150 EH_TC_STAGGER_NS=150
run s300 EH_TC_STAGGER_NS=300
run s500 EH_TC_STAGGER_NS=500
run s800 EH_TC_STAGGER_NS=800
