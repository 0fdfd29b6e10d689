#!/bin/bash
export PYTHONFAULTHANDLER=1
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "torchrun rc=$?"
tail -c 600 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_n$N.json'))
print('N=$N value', d['value'], 'us/step', d['ms_per_step']*1e3, 'e2e', d['e2e']['value'], 'dp_parity', d.get('dp_parity_ok'))
print('wide', {k:v for k,v in d['extra']['c5_wide_mlp'].items() if k in('value','us_per_step','n_gpus','error')})
print('strong', d['extra']['c4_strong_scaling'])
PY
