#!/bin/bash
run() { # name, env...
  name=$1; shift
  env "$@" timeout 120 python bench.py --steps 512 --warmup 32 --no-cpu-baseline --no-wide --no-e2e > gpurun_out/r2_j14_$name.json 2>> gpurun_out/r2_j14.err
  python -c "
import json
d=json.load(open('gpurun_out/r2_j14_$name.json')); print('$name', 'us/step', round(d['ms_per_step']*1e3,3))"
  env "$@" EH_EPOCH_DEBUG=gpurun_out/r2_j14_$name.bin EH_PROF_LOG2N=24 timeout 120 python tools/epoch_prof_driver.py 0 32 > /dev/null 2>&1
  python tools/epoch_phase_dump.py gpurun_out/r2_j14_$name.bin 2>&1 | sed -n '10,19p'
}
run tc128 A=1
run tc148 EH_EPOCH_GRID=148
run ffma147 EH_NO_TC=1
run ffma128 EH_NO_TC=1 EH_EPOCH_GRID=128
run ffma128w15 EH_NO_TC=1 EH_EPOCH_GRID=128 EH_EPOCH_WARPS=15
