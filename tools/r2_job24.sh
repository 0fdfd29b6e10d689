#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -x -q -m gpu -k "rbq10 or c3 or tensor_engine or determinism or persistent" --timeout=300 2>&1 | tail -4
bash tools/r2_job22.sh
