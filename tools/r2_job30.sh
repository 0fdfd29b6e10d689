#!/bin/bash
# memcheck over the exact-fp32 part of the GPU suite (the wide tests take minutes under the sanitizer)
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_generic.py tests/test_abi_replay.py -x -q -m gpu --timeout=600 -k "not test_train_api" > gpurun_out/r2_memcheck_gpu_suite.txt 2>&1; echo "memcheck rc=$?"
grep -E "passed|failed|ERROR SUMMARY|Invalid|leak" gpurun_out/r2_memcheck_gpu_suite.txt | head -12
