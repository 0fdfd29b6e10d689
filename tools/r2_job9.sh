#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "host or stream or zero_copy" 2>&1 | tail -5
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-wide > gpurun_out/r2_j9_bench.json 2> gpurun_out/r2_j9_bench.err; tail -c 300 gpurun_out/r2_j9_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2_j9_bench.json')); print('value', d['value'], 'us/step', d['ms_per_step']*1e3, 'e2e', d['e2e']['value'], d['e2e']['resident_dataset']['value'])"
python bench.py --steps 2048 --warmup 64 --no-cpu-baseline --no-wide > gpurun_out/r2_j9_bench_long.json 2>> gpurun_out/r2_j9_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2_j9_bench_long.json')); print('LONG value', d['value'], 'us/step', d['ms_per_step']*1e3, 'e2e', d['e2e']['value'], d['e2e']['resident_dataset']['value'])"
