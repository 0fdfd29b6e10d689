#!/bin/bash
export PYTHONUNBUFFERED=1 PYTHONFAULTHANDLER=1
timeout 600 python -u -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 --log-dir gpurun_out/r2_j21_logs --tee 3 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_j21_n2.out 2> gpurun_out/r2_j21_n2.err; echo "torchrun rc=$?"
tail -c 1200 gpurun_out/r2_j21_n2.err; echo; head -c 600 gpurun_out/r2_j21_n2.out
find gpurun_out/r2_j21_logs -type f | head; for f in $(find gpurun_out/r2_j21_logs -name "*.log" | head -4); do echo "== $f"; tail -c 800 $f; done
