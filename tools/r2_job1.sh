#!/bin/bash
# round 2, job 1: stall-reason profiles of the per-chunk code (k_step, both engines) + baseline bench lines
set -x
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-wide > gpurun_out/r2_base_bench.json 2> gpurun_out/r2_base_bench.err
EH_CLUSTER_SIZE=2 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-wide --no-e2e > gpurun_out/r2_base_bench_cs2.json 2>> gpurun_out/r2_base_bench.err
EH_CLUSTER_SIZE=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-wide --no-e2e > gpurun_out/r2_base_bench_cs1.json 2>> gpurun_out/r2_base_bench.err
ncu --set full --clock-control none --import-source on -k regex:k_step -s 4 -c 1 -o gpurun_out/r2_k_step_ffma -f python tools/epoch_prof_driver.py 7 8 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step -s 4 -c 1 -o gpurun_out/r2_k_step_mma -f python tools/epoch_prof_driver.py 23 8 > gpurun_out/ncu2.log 2>&1
EH_CLUSTER_SIZE=1 ncu --set full --clock-control none --import-source on -k regex:k_epoch -c 1 -o gpurun_out/r2_k_epoch_cs1 -f python tools/epoch_prof_driver.py 0 12 > gpurun_out/ncu3.log 2>&1
echo "ncu coop cs1 rc=$?"
tail -2 gpurun_out/ncu3.log
ls -la gpurun_out/
