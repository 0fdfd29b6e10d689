#!/bin/bash
cp easyhybrid.jl_b200/libeasyhybrid_cuda.so /tmp/orig.so
for v in orig r2 r4 r16 r32 orig; do
if [ $v = orig ]; then cp /tmp/orig.so easyhybrid.jl_b200/libeasyhybrid_cuda.so; else cp easyhybrid.jl_b200/libeh_$v.so easyhybrid.jl_b200/libeasyhybrid_cuda.so; fi
timeout 120 python bench.py --steps 1024 --warmup 32 --no-cpu-baseline --no-wide --no-e2e | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', 'us/step', round(d['ms_per_step']*1e3,3), d['final_loss'])"
done
cp /tmp/orig.so easyhybrid.jl_b200/libeasyhybrid_cuda.so
