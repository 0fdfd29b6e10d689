#!/bin/bash
# round-end evidence: full GPU test suite, the bench line, the ncu launch list of the same bench command (short), and one
# full ncu capture of the persistent kernel
python -m pytest tests -q -m gpu 2>&1 | tail -4
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 600 gpurun_out/bench_final.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_final.csv \
    python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches_bench_final.csv
EH_NO_COOP=1 ncu --set full --clock-control none --import-source on -k regex:k_epoch -c 1 -o gpurun_out/k_epoch_final -f \
    python tools/epoch_prof_driver.py 0 12 > gpurun_out/ncu_final.log 2>&1
tail -3 gpurun_out/ncu_final.log
