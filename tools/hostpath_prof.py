"""Where does the time of a host-batch step go?  Enqueue time (CPU) vs total time, for several batch sizes."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import easyhybrid_b200 as eh
from bench import make_model, synth

model = make_model(eh)
for B in (65536, 16384, 4096):
    n = 16 * B
    xf, y = synth(n, 1)
    sess = eh.FusedSession(model, opt=eh.Adam(0.01), device=0)
    sess.upload(0, xf, y)
    sess.set_params(model.initialparameters(np.random.default_rng(0)))
    hb = []
    for i in range(16):
        sl = slice(i * B, (i + 1) * B)
        hb.append(sess.host_batch(sess.pinned(xf[0][sl]), [sess.pinned(xf[1]["ta"][sl])], [sess.pinned(y["reco"][sl])]))
    K = 2048
    el = sess.pinned(np.zeros(K + 8, dtype=np.float32))
    for i in range(64):
        sess.step_host_async(hb[i % 16], el, i % 8)
    sess.sync()
    t0 = time.perf_counter()
    for i in range(K):
        sess.step_host_async(hb[i % 16], el, i)
    t1 = time.perf_counter()
    sess.sync()
    t2 = time.perf_counter()
    print(f"B={B}: enqueue {1e6*(t1-t0)/K:.1f} us/step, total {1e6*(t2-t0)/K:.1f} us/step, zero_copy={os.environ.get('EH_HOST_NO_ZEROCOPY','0')!='1'}", flush=True)
    sess.close()
