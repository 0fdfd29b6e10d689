#!/bin/bash
export EH_TC_MIN_BATCH=16384
EH_DEBUG_GEOM=1 EH_EPOCH_DEBUG=gpurun_out/r2_j18.bin EH_PROF_LOG2N=24 timeout 120 python tools/epoch_prof_driver.py 0 8 > gpurun_out/r2_j18.txt 2>&1
grep -E "DBG|Error|\[eh\]" gpurun_out/r2_j18.txt | head
