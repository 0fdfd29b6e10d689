#!/bin/bash
# host-batch path: parity tests, then the bench with and without the zero-copy packer
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "host" 2>&1 | tail -5
python bench.py --no-wide --no-cpu-baseline > gpurun_out/b13.json 2> gpurun_out/b13.err
EH_HOST_NO_ZEROCOPY=1 python bench.py --no-wide --no-cpu-baseline > gpurun_out/b13n.json 2> gpurun_out/b13n.err
python - <<'PY'
import json
for f in ("b13", "b13n"):
    try:
        d = json.loads(open("gpurun_out/" + f + ".json").read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d["e2e"]["resident_dataset"]["value"], d["clocks"])
    except Exception as e:
        print(f, "failed", e, open("gpurun_out/" + f + ".err").read()[-2000:])
PY
