#!/bin/bash
export EH_TC_MIN_BATCH=16384
cp easyhybrid.jl_b200/libeasyhybrid_cuda.so /tmp/orig.so
for v in A B; do
cp easyhybrid.jl_b200/libeh_$v.so easyhybrid.jl_b200/libeasyhybrid_cuda.so
for i in 1 2 3; do
EH_PROF_LOG2N=24 timeout 120 python tools/epoch_prof_driver.py 0 16 > gpurun_out/r2_j19_$v$i.txt 2>&1; echo "variant $v run $i: $(tail -1 gpurun_out/r2_j19_$v$i.txt | cut -c1-150)"
done
done
cp /tmp/orig.so easyhybrid.jl_b200/libeasyhybrid_cuda.so
