#!/bin/bash
timeout 60 tools/_bin/tc_proto > gpurun_out/r2_tc_proto.txt 2>&1; tail -24 gpurun_out/r2_tc_proto.txt
timeout 600 python -m pytest tests/test_gpu_baseline_sizes.py -x -q -m gpu -k "c3" 2>&1 | tail -15
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-wide --no-e2e > gpurun_out/r2_j11_bench.json 2> gpurun_out/r2_j11_bench.err; tail -c 400 gpurun_out/r2_j11_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2_j11_bench.json')); print('value', d['value'], 'us/step', d['ms_per_step']*1e3, 'loss', d.get('final_loss'))"
EH_NO_TC=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-wide --no-e2e > gpurun_out/r2_j11_bench_notc.json 2>> gpurun_out/r2_j11_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2_j11_bench_notc.json')); print('NO_TC value', d['value'], 'us/step', d['ms_per_step']*1e3, 'loss', d.get('final_loss'))"
EH_EPOCH_DEBUG=gpurun_out/r2_j11_dbg.bin EH_PROF_LOG2N=24 timeout 120 python tools/epoch_prof_driver.py 0 32 > gpurun_out/r2_j11_dbg.log 2>&1
python tools/epoch_phase_dump.py gpurun_out/r2_j11_dbg.bin > gpurun_out/r2_j11_phases.txt 2>&1; head -30 gpurun_out/r2_j11_phases.txt
