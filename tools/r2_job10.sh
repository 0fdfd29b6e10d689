#!/bin/bash
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_j10_bench.json 2> gpurun_out/r2_j10_bench.err; tail -c 300 gpurun_out/r2_j10_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2_j10_bench.json')); print('value', d['value'], 'us/step', d['ms_per_step']*1e3, 'e2e', d['e2e']['value'], d['e2e']['resident_dataset']['value'], 'wide', d['extra']['c5_wide_mlp'].get('us_per_step'), 'cpu', d['cpu_baseline']['value'])"
