#!/bin/bash
set -x
python tools/parity_table.py rbq10 > gpurun_out/r2_parity_table_fast_tanh.txt 2>&1; tail -32 gpurun_out/r2_parity_table_fast_tanh.txt
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-wide --no-e2e > gpurun_out/r2_j5_bench.json 2> gpurun_out/r2_j5_bench.err; tail -c 300 gpurun_out/r2_j5_bench.err
python bench.py --steps 2048 --warmup 64 --no-cpu-baseline --no-wide --no-e2e > gpurun_out/r2_j5_bench_long.json 2>> gpurun_out/r2_j5_bench.err
EH_USE_X2=1 python bench.py --steps 2048 --warmup 64 --no-cpu-baseline --no-wide --no-e2e > gpurun_out/r2_j5_bench_long_x2.json 2>> gpurun_out/r2_j5_bench.err
python bench.py --steps 2048 --warmup 64 --no-cpu-baseline --no-wide --no-e2e --flags 16 > gpurun_out/r2_j5_bench_long_mma.json 2>> gpurun_out/r2_j5_bench.err
EH_EPOCH_DEBUG=gpurun_out/r2_j5_dbg.bin EH_PROF_LOG2N=24 python tools/epoch_prof_driver.py 0 32 > gpurun_out/r2_j5_dbg.log 2>&1
python tools/epoch_phase_dump.py gpurun_out/r2_j5_dbg.bin > gpurun_out/r2_j5_phases.txt 2>&1
head -14 gpurun_out/r2_j5_phases.txt
for f in gpurun_out/r2_j5_bench*.json; do python -c "
import json
d=json.load(open('$f')); print('$f', d['value'], d['ms_per_step']*1e3)"; done
