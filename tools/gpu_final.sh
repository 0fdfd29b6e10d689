#!/bin/bash
# Round-end evidence on one B200 (run under gpurun): the GPU test suite, the bench lines, and the captures profiles/ is
# built from (tools/ncu_metrics_to_json.py, tools/epoch_phase_dump.py, tools/sass_summary.py turn them into the committed
# summaries).  Nothing printed under ncu is a bench value.
python -m pytest tests -q -m gpu --timeout=300 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 400 gpurun_out/bench_n1.err
python bench.py --steps 2048 --warmup 64 --no-cpu-baseline --no-wide > gpurun_out/bench_n1_long.json 2>> gpurun_out/bench_n1.err
EH_TC_MIN_BATCH=16384 python bench.py --steps 2048 --warmup 64 --no-cpu-baseline --no-wide --no-e2e > gpurun_out/bench_n1_long_tc.json 2>> gpurun_out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_n1.err
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
# full captures of the persistent kernel (FFMA2 engine, tensor engine) and of the wide GEMMs
ncu --set full --clock-control none --import-source on -k regex:k_epoch -c 1 -o gpurun_out/k_epoch -f \
    python tools/epoch_prof_driver.py 0 12 > gpurun_out/ncu_1.log 2>&1
EH_TC_MIN_BATCH=16384 ncu --set full --clock-control none --import-source on -k regex:k_epoch -c 1 -o gpurun_out/k_epoch_tc -f \
    python tools/epoch_prof_driver.py 0 12 > gpurun_out/ncu_2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wide_gemm -s 40 -c 3 -o gpurun_out/wide_gemm -f \
    python tools/wide_bench.py > gpurun_out/ncu_3.log 2>&1
# per-phase SM-clock stamps of the persistent kernel
EH_EPOCH_DEBUG=gpurun_out/phases_ffma.bin EH_PROF_LOG2N=24 python tools/epoch_prof_driver.py 0 32 > /dev/null 2>&1
python tools/epoch_phase_dump.py gpurun_out/phases_ffma.bin > gpurun_out/phases_ffma.txt 2>&1
EH_TC_MIN_BATCH=16384 EH_EPOCH_DEBUG=gpurun_out/phases_tc.bin EH_PROF_LOG2N=24 python tools/epoch_prof_driver.py 0 32 > /dev/null 2>&1
python tools/epoch_phase_dump.py gpurun_out/phases_tc.bin > gpurun_out/phases_tc.txt 2>&1
# multi-GPU (gpurun --gpus N): python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
#     --master-port 29521 bench.py --gpus N --steps 20 --warmup 5 --no-cpu-baseline
