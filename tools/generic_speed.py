"""us per step of a traced process model -- interpreted per sample (generic variant) and compiled at run time (NVRTC) -- next to
the specialised RbQ10 variant, resident data"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import easyhybrid_b200 as eh
from bench import make_model, synth


def custom(*, ta, rb, Q10, tref=15.0):
    return {"reco": rb * Q10 ** (0.1 * (ta - tref)) + 0.0 * ta}


n = 1 << 22
xf, y = synth(n, 1)
models = {
    "specialised RbQ10": make_model(eh),
    "generic (traced RbQ10 + 0*ta)": eh.constructHybridModel(["sw_pot", "dsw_pot"], ["ta"], ["reco"], custom,
                                                             dict(rb=(3.0, 0.0, 13.0), Q10=(2.0, 1.0, 4.0)), ["rb"], ["Q10"],
                                                             hidden_layers=[16, 16], activation="tanh", scale_nn_outputs=True),
}
models["compiled (NVRTC) traced RbQ10 + 0*ta"] = models["generic (traced RbQ10 + 0*ta)"]
for name, model in models.items():
    for B in (512, 65536):
        sess = eh.FusedSession(model, opt=eh.Adam(0.01), device=0, jit=name.startswith("compiled"))
        sess.upload(0, xf, y)
        sess.set_params(model.initialparameters(np.random.default_rng(0)))
        sess.set_perm(np.random.default_rng(7).permutation(n))
        sess.run_steps(B, 0, 64)
        K = 1024
        losses = sess.run_steps(B, 64, K)
        ms, _, _ = sess.last_timing()
        print(f"{name:40s} {sess.kernel_variant():55s} B={B:6d}: {1e3*ms/K:7.2f} us/step  loss {losses[-1]:.4f}", flush=True)
        sess.close()
