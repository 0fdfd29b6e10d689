"""Decode the EH_EPOCH_DEBUG timestamp dump of the persistent kernel (SM clocks per phase; clocks are per SM, so
only differences inside one CTA are meaningful).

Stamps of thread 0 of every CTA: 0 step top, 1 its compute phase + row hand-over done, 2 CTA barrier issued,
3 partial published (A; includes the wait for the CTA's slowest warp -- BAR.SYNC blocks at the next dependent
instruction, not at issue), 4 totals in shared memory (B somewhere + C), 5 optimiser done, 6 end-of-step barrier issued.
8 + w: end of warp w's compute phase.  Service warp: 24 / 25 start / end of its slice reduction (B), 26 totals gathered (C),
27 global-parameter tail published."""
import sys
import numpy as np
raw = np.fromfile(sys.argv[1], dtype=np.int64)
nsteps, G, w, tiles = raw[:4]
d = raw[4:].reshape(nsteps, G, 32)
names = ["compute (warp 0) + rows", "CTA barrier (issue only)", "A: wait slowest warp, sum rows, publish", "B + C: until the totals are in smem",
         "optimiser (theta entry)", "next scalars"]
print(f"steps {nsteps} grid {G} compute warps {w} tiles {tiles}")
for s in range(1, nsteps):
    if s not in (1, nsteps // 2, nsteps - 1):
        continue
    t = d[s, :, :7].astype(np.float64)
    ph = np.diff(t, axis=1)
    nxt = d[s + 1, :, 0] - d[s, :, 0] if s + 1 < nsteps else t[:, 6] - t[:, 0]
    print(f"step {s}: step-top to step-top cycles median {np.median(nxt):.0f} max {nxt.max():.0f}")
    for i, nme in enumerate(names):
        print(f"   {nme:42s} median {np.median(ph[:, i]):8.0f}  min {ph[:, i].min():8.0f}  max {ph[:, i].max():8.0f}")
    wend = d[s, :, 8:8 + w].astype(np.float64) - t[:, 0:1]
    wend = np.where(d[s, :, 8:8 + w] > 0, wend, np.nan)
    print("   per-warp compute end (median over CTAs): " + " ".join(f"{np.nanmedian(wend[:, i]):.0f}" for i in range(w)))
    own = d[s, :, 24] > 0
    if own.any():
        b0 = (d[s, own, 24] - d[s, own, 0]).astype(np.float64)
        b1 = (d[s, own, 25] - d[s, own, 0]).astype(np.float64)
        c1 = (d[s, own, 26] - d[s, own, 0]).astype(np.float64)
        print(f"   service warp (since step top): B starts {np.median(b0):.0f}, B done {np.median(b1):.0f} (max {b1.max():.0f}), totals gathered {np.median(c1):.0f} (max {c1.max():.0f})")
    if d[s, :, 27].any():
        tail = (d[s, :, 27] - d[s, :, 0]).astype(np.float64)
        print(f"   service warp: global-parameter tail published {np.median(tail):.0f} after the step top")
