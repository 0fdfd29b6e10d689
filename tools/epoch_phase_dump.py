"""Decode the EH_EPOCH_DEBUG timestamp dump of the persistent kernel (SM clocks per phase)."""
import sys
import numpy as np
raw = np.fromfile(sys.argv[1], dtype=np.int64)
nsteps, G, w, cs = raw[:4]
d = raw[4:].reshape(nsteps, G, 32)
names = ["scalars+sync", "compute(thread0 warp)", "cta_reduce", "hop1 cluster->leader->L2", "hop2 L2 share sum + push", "hop3 cluster shares", "dp exchange+adam+init"]
print(f"steps {nsteps} grid {G} warps {w} cluster {cs}")
for s in range(1, nsteps):
    t = d[s, :, :8].astype(np.float64)
    ph = np.diff(t, axis=1)                       # 7 phases per CTA
    wend = d[s, :, 8:8 + w].astype(np.float64) - t[:, 1:2]   # per-warp compute time since phase-1 stamp
    tot = (d[s, :, 7] - d[s, :, 0]).astype(np.float64)
    if s in (1, nsteps // 2, nsteps - 1):
        print(f"step {s}: total cycles median {np.median(tot):.0f} max {tot.max():.0f}")
        for i, nme in enumerate(names):
            print(f"   {nme:28s} median {np.median(ph[:, i]):8.0f}  min {ph[:, i].min():8.0f}  max {ph[:, i].max():8.0f}")
        if d[s, :, 27].any():
            rd = (d[s, :, 27] - d[s, :, 6]).astype(np.float64); ex = (d[s, :, 28] - d[s, :, 27]).astype(np.float64); ad = (d[s, :, 7] - d[s, :, 28]).astype(np.float64)
            print(f"     of which: rank exchange median {np.median(ex):.0f} (max {ex.max():.0f}), adam+init median {np.median(ad):.0f}")
        if d[s, :, 29].any() and d[s, :, 30].any():
            # finer stamps of thread 0 (every stamp itself costs ~350 cycles: clock read + global store)
            sc = (d[s, :, 29] - d[s, :, 28]).astype(np.float64); lo = (d[s, :, 30] - d[s, :, 29]).astype(np.float64)
            print(f"     adam phase of thread 0: batch scalars median {np.median(sc):.0f}, its one parameter {np.median(lo):.0f}"
                  " (the step-top barrier then waits for the thread that owns a phi entry: sigmoid, squashing, double log2)")
        print(f"   per-warp compute: median {np.median(wend):.0f} min {wend.min():.0f} max {wend.max():.0f}; per-CTA slowest warp median {np.median(wend.max(axis=1)):.0f}")
