#!/bin/bash
export EH_TC_MIN_BATCH=16384
EH_PROF_LOG2N=22 timeout 300 compute-sanitizer --tool racecheck --print-limit 12 python tools/epoch_prof_driver.py 0 4 > gpurun_out/r2_j17_racecheck.txt 2>&1
grep -v "Host Frame\|Saved host\|^========= $" gpurun_out/r2_j17_racecheck.txt | head -60
EH_PROF_LOG2N=22 timeout 300 compute-sanitizer --tool initcheck --print-limit 8 python tools/epoch_prof_driver.py 0 4 > gpurun_out/r2_j17_initcheck.txt 2>&1
grep -v "Host Frame\|Saved host\|^========= $" gpurun_out/r2_j17_initcheck.txt | head -30
