#!/bin/bash
set -x
ncu --set full --clock-control none --import-source on -k regex:k_epoch -c 1 -o gpurun_out/r2_k_epoch_v8 -f python tools/epoch_prof_driver.py 0 12 > gpurun_out/ncu_a.log 2>&1
EH_USE_X2=1 ncu --set full --clock-control none --import-source on -k regex:k_epoch -c 1 -o gpurun_out/r2_k_epoch_v8_x2 -f python tools/epoch_prof_driver.py 0 12 > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_epoch -c 1 -o gpurun_out/r2_k_epoch_v8_mma -f python tools/epoch_prof_driver.py 16 12 > gpurun_out/ncu_c.log 2>&1
tail -2 gpurun_out/ncu_a.log gpurun_out/ncu_b.log gpurun_out/ncu_c.log
