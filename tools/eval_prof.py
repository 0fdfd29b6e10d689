"""eval kernel (K5) throughput: test-mode forward + metric statistics over a resident split"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import easyhybrid_b200 as eh
from bench import make_model, synth

model = make_model(eh)
for n in (1 << 20, 1 << 24):
    xf, y = synth(n, 1)
    sess = eh.FusedSession(model, opt=eh.Adam(0.01), device=0)
    sess.upload(0, xf, y)
    sess.set_params(model.initialparameters(np.random.default_rng(0)))
    for want in (False, True):
        sess.eval(0, want_yhat=want)
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            sess.eval(0, want_yhat=want)
        dt = (time.perf_counter() - t0) / reps
        ms, _, _ = sess.last_timing()
        print(f"n={n} want_yhat={want}: kernel {ms*1e3:.1f} us = {n*16/ms/1e6:.1f} GB/s algorithmic, {n/ms/1e6:.2f} Gsamples/s; call wall {dt*1e3:.2f} ms", flush=True)
    sess.close()
