#!/bin/bash
export EH_TC_MIN_BATCH=16384
for i in 1 2 3 4; do
EH_PROF_LOG2N=22 timeout 120 python tools/epoch_prof_driver.py 0 32 2>&1 | tail -1
done
EH_PROF_LOG2N=22 timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/epoch_prof_driver.py 0 8 > gpurun_out/r2_j16_memcheck.txt 2>&1; grep -v "^=========     Host Frame\|^=========         in " gpurun_out/r2_j16_memcheck.txt | head -40
