"""persistent-kernel geometry vs batch size: us per step for EH_EPOCH_WARPS settings (resident data)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import easyhybrid_b200 as eh
from bench import make_model, synth

model = make_model(eh)
n = 1 << 20
xf, y = synth(n, 1)
for B in (12, 256, 512, 1024, 4096):
    row = []
    for w in ("",):
        if w:
            os.environ["EH_EPOCH_WARPS"] = w
        else:
            os.environ.pop("EH_EPOCH_WARPS", None)
        sess = eh.FusedSession(model, opt=eh.Adam(0.01), device=0)
        sess.upload(0, xf, y)
        sess.set_params(model.initialparameters(np.random.default_rng(0)))
        sess.set_perm(np.random.default_rng(7).permutation(n))
        sess.run_steps(B, 0, 64)
        K = 1024
        sess.run_steps(B, 64, K)
        ms, launches, _ = sess.last_timing()
        row.append(f"w={w or 'auto'}: {1e3*ms/K:.2f}")
        sess.close()
    print(f"B={B}: " + "  ".join(row) + "  us/step", flush=True)
