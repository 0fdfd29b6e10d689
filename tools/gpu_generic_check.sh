#!/bin/bash
python -m pytest tests/test_gpu_generic.py tests/test_gpu_parity.py -q -m gpu -k "generic or traced or zero_copy" -x 2>&1 | tail -40 > gpurun_out/generic.log
tail -15 gpurun_out/generic.log
