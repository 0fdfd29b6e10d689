#!/bin/bash
ncu --set full --clock-control none --import-source on -k regex:k_epoch -c 1 -o gpurun_out/r2_k_epoch_tc1 -f python tools/epoch_prof_driver.py 0 12 > gpurun_out/ncu_tc1.log 2>&1
tail -2 gpurun_out/ncu_tc1.log
