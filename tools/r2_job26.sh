#!/bin/bash
ncu --set full --clock-control none --import-source on -k regex:k_epoch -c 1 -o gpurun_out/r2_k_epoch_final -f python tools/epoch_prof_driver.py 0 12 > gpurun_out/ncu_f1.log 2>&1; tail -1 gpurun_out/ncu_f1.log
EH_TC_MIN_BATCH=16384 ncu --set full --clock-control none --import-source on -k regex:k_epoch -c 1 -o gpurun_out/r2_k_epoch_tc_final -f python tools/epoch_prof_driver.py 0 12 > gpurun_out/ncu_f2.log 2>&1; tail -1 gpurun_out/ncu_f2.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_f3.log 2>&1; tail -c 300 gpurun_out/ncu_f3.log
ncu --set full --clock-control none --import-source on -k regex:k_wide_gemm -s 40 -c 3 -o gpurun_out/r2_wide_gemm_final -f python tools/wide_bench.py > gpurun_out/ncu_f4.log 2>&1; tail -2 gpurun_out/ncu_f4.log
EH_EPOCH_DEBUG=gpurun_out/r2_final_ffma.bin EH_PROF_LOG2N=24 python tools/epoch_prof_driver.py 0 32 > /dev/null 2>&1; python tools/epoch_phase_dump.py gpurun_out/r2_final_ffma.bin > gpurun_out/r2_final_phases_ffma.txt 2>&1
EH_TC_MIN_BATCH=16384 EH_EPOCH_DEBUG=gpurun_out/r2_final_tc.bin EH_PROF_LOG2N=24 python tools/epoch_prof_driver.py 0 32 > /dev/null 2>&1; python tools/epoch_phase_dump.py gpurun_out/r2_final_tc.bin > gpurun_out/r2_final_phases_tc.txt 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err; tail -c 300 gpurun_out/r2_final_bench_n1.err
python bench.py --steps 2048 --warmup 64 --no-cpu-baseline --no-wide > gpurun_out/r2_final_bench_n1_long.json 2>> gpurun_out/r2_final_bench_n1.err
EH_TC_MIN_BATCH=16384 python bench.py --steps 2048 --warmup 64 --no-cpu-baseline --no-wide --no-e2e > gpurun_out/r2_final_bench_n1_long_tc.json 2>> gpurun_out/r2_final_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_bench_ref.json 2>> gpurun_out/r2_final_bench_n1.err
for f in r2_final_bench_n1 r2_final_bench_n1_long r2_final_bench_n1_long_tc; do python -c "
import json
d=json.load(open('gpurun_out/$f.json')); e=d.get('e2e') or {}; print('$f', d['steps'], 'us/step', round(d['ms_per_step']*1e3,3), 'value', d['value'], 'e2e', e.get('value'), (e.get('resident_dataset') or {}).get('value'), 'wide', ((d.get('extra') or {}).get('c5_wide_mlp') or {}).get('us_per_step'))"; done
