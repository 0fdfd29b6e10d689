#!/bin/bash
export EH_BENCH_DEBUG=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_dbg_n2.json 2> gpurun_out/r2_dbg_n2.err; echo "torchrun rc=$?"
grep "bench\]" gpurun_out/r2_dbg_n2.err; ls -la gpurun_out/r2_dbg_n2.json; head -c 300 gpurun_out/r2_dbg_n2.json
