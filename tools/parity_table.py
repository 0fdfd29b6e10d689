"""Prints the loss / gradient errors of the GPU path against the float64 oracle for the RbQ10 cases of
tests/test_gpu_parity.py (no assertions): used to judge numerics changes (e.g. the tanh form)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import easyhybrid_b200 as eh
from oracle import oracle as orc
import conftest as cf
from test_gpu_parity import CASES, _setup
only = sys.argv[1] if len(sys.argv) > 1 else "rbq10"
worst = 0
for (name, mk, mkdata, loss, agg) in CASES:
    if only not in name and only != "all":
        continue
    for flags in (0, 16):
        model, xf, y, flat, sess, o, rng = _setup(eh, orc, mk, mkdata, loss, agg, flags=flags)
        n = xf[0].shape[0]
        for B in (n, 517, 64, 12, 1):
            if B == 1 and loss in ("nseLoss", "PT"):
                continue
            idx = rng.permutation(n)[:B]
            if np.isnan(np.stack([y[t][idx] for t in model.targets])).all():
                continue
            L, g = sess.loss_grad(idx)
            L64, g64 = o.loss_grad(flat, xf, y, idx, precision=64)
            L32, g32 = o.loss_grad(flat, xf, y, idx, precision=32)
            sc = np.abs(g64).max()
            e, e32 = np.abs(g - g64).max() / sc, np.abs(g32 - g64).max() / sc
            worst = max(worst, e)
            print(f"{name:22s} flags {flags:2d} B {B:5d}  loss rel {abs(L - L64) / abs(L64):.2e} (f32 oracle {abs(L32 - L64) / abs(L64):.2e})  grad {e:.2e} (f32 oracle {e32:.2e}) {'  <-- > 1e-5' if e > 1e-5 else ''}")
        sess.close()
print("worst gradient error", worst)
