#!/bin/bash
# round 2, job 2: new persistent kernel (reduce-scatter/all-gather exchange, service warp, TMA record tiles, epoch staging)
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-wide > gpurun_out/r2_j2_bench.json 2> gpurun_out/r2_j2_bench.err; tail -c 300 gpurun_out/r2_j2_bench.err
EH_NO_STAGE=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-wide --no-e2e > gpurun_out/r2_j2_bench_nostage.json 2>> gpurun_out/r2_j2_bench.err
python bench.py --steps 2048 --warmup 64 --no-cpu-baseline --no-wide --no-e2e > gpurun_out/r2_j2_bench_long.json 2>> gpurun_out/r2_j2_bench.err
EH_EPOCH_DEBUG=gpurun_out/r2_j2_dbg.bin EH_PROF_LOG2N=24 python tools/epoch_prof_driver.py 0 32 > gpurun_out/r2_j2_dbg.log 2>&1
python tools/epoch_phase_dump.py gpurun_out/r2_j2_dbg.bin > gpurun_out/r2_j2_phases.txt 2>&1
cat gpurun_out/r2_j2_phases.txt
EH_NO_STAGE=1 EH_EPOCH_DEBUG=gpurun_out/r2_j2_dbg_ns.bin EH_PROF_LOG2N=24 python tools/epoch_prof_driver.py 0 32 > gpurun_out/r2_j2_dbg_ns.log 2>&1
python tools/epoch_phase_dump.py gpurun_out/r2_j2_dbg_ns.bin > gpurun_out/r2_j2_phases_nostage.txt 2>&1
for f in gpurun_out/r2_j2_bench*.json; do python -c "
import json
d=json.load(open('$f')); print('$f', d['value'], d['ms_per_step']*1e3, d.get('e2e') and d['e2e']['value'], d.get('e2e') and d['e2e'].get('resident_dataset',{}).get('value'))"; done
