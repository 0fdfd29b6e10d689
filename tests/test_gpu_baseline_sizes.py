"""Parity against the oracle at the BASELINE configurations' OWN sizes (BASELINE.json configs):

  C3  RbQ10 [2-16-16-1] tanh, mse, Adam(0.01), batch 65 536 drawn from N >= 2^20 samples
  C2  Linear_Regression through the generic model: NN [15,15] relu -> a, global b, Adam(0.001), batch 100
      (projects/Linear_Regression/linearRegression.jl:13-18 via test/test_generic_hybrid_model.jl:10-20)
  C5  two-target Expo hybrid, hidden [512,512,512], PerTarget(nseLoss, mse), batch 65 536 (bf16 tcgen05 path)

The truth is the float64 oracle.  For the fp32 trajectories the reference trajectory is built step by step from the
float64 loss / gradient of the oracle and the oracle's own (Float32, Optimisers.jl semantics) optimiser rule, so
every step of the GPU run is compared with a value that carries no float32 forward/backward error of its own.

Bounds and why (north_star: loss and gradient within 1e-5 relative in fp32, phi within 1e-4 after a fixed number of steps):
  * single loss / gradient evaluations: 1e-5 (loss relative; gradient relative to its max-norm), no exceptions;
  * trajectories: the per-step loss is asserted at 1e-5 for EVERY step.  That is possible because the steps compared
    start from the same parameters only at step 0; later steps carry the divergence of two Float32 optimiser
    trajectories (Adam turns a noise-level gradient entry into a +-eta step, src of the caveat: SURVEY 10.5), which on
    these configurations stays below 1e-5 in the loss for the 50 steps checked (measured: see the assert messages);
  * C5 (bf16 tensor cores): stated tolerance 2e-2 on the loss, and an element-wise bound per parameter block:
    max |g - g64| <= 4e-2 of the block's max-norm (next to the direction / norm checks of tests/test_gpu_wide.py)."""
import numpy as np
import pytest

from conftest import linear_model, make_expo2, make_linear, make_synth, rbq10_model, truth_trajectory as _truth_trajectory

pytestmark = pytest.mark.gpu

B_C3 = 65536


def _q10(p):
    return 1.0 + 3.0 / (1.0 + np.exp(-float(p[-1])))


def test_c3_loss_and_gradient_batch_65536(eh, orc):
    """BASELINE config 3 at its own batch size: one fused step's loss and gradient vs the float64 oracle"""
    model = rbq10_model(eh)
    n = 1 << 20
    xf, y = eh.prepare_data(model, make_synth(n))
    rng = np.random.default_rng(11)
    flat = model.initialparameters(rng)
    flat += (0.05 * rng.standard_normal(flat.size)).astype(np.float32)
    o = orc.Oracle(model, opt=eh.Adam(0.01))
    nt = orc.max_threads()
    for flags in (0, 4, 16):   # persistent-kernel engine, step-kernel pair, tensor-pipe engine
        sess = eh.FusedSession(model, opt=eh.Adam(0.01), flags=flags)
        sess.upload(0, xf, y)
        sess.set_params(flat)
        for seed in (1, 2):
            idx = np.random.default_rng(seed).permutation(n)[:B_C3]
            L, g = sess.loss_grad(idx)
            L64, g64 = o.loss_grad(flat, xf, y, idx, precision=64, nthreads=nt)
            assert abs(L - L64) <= 1e-5 * abs(L64), (flags, L, L64)
            err = np.abs(g - g64).max() / np.abs(g64).max()
            assert err <= 1e-5, (flags, err)
        sess.close()


def test_c3_trajectory_50_steps_batch_65536(eh, orc):
    """50 Adam(0.01) steps of batch 65 536 (persistent kernel, one launch): EVERY per-step loss within 1e-5 of the
    float64-gradient trajectory, Q10 within 1e-4 relative after the 50 steps (north_star)"""
    model = rbq10_model(eh)
    n = 1 << 22
    xf, y = eh.prepare_data(model, make_synth(n))
    flat = model.initialparameters(np.random.default_rng(12))
    perm = np.random.default_rng(13).permutation(n)[: 50 * B_C3]
    o = orc.Oracle(model, opt=eh.Adam(0.01))
    want, ref = _truth_trajectory(o, flat, xf, y, perm, B_C3, orc.max_threads())
    sess = eh.FusedSession(model, opt=eh.Adam(0.01))
    sess.upload(0, xf, y)
    sess.set_params(flat)
    got = sess.epoch(perm, B_C3)
    rel = np.abs(got - want) / np.abs(want)
    assert got.shape == (50,) and rel.max() <= 1e-5, (int(rel.argmax()), float(rel.max()), rel[:5])
    ps = sess.get_params()
    assert abs(_q10(ps) - _q10(ref)) <= 1e-4 * _q10(ref), (_q10(ps), _q10(ref))
    assert want[-1] < 0.5 * want[0]          # and it is a training run, not a fixed point
    sess.close()


def test_c2_linear_adam_1e3_batch_100_trajectory(eh, orc):
    """BASELINE config 2, Linear_Regression at its own settings: [15,15] relu, Adam(0.001), batch 100, N = 1000:
    five epochs (50 steps, fresh permutation each) vs the float64-gradient trajectory"""
    model = linear_model(eh)
    xf, y = eh.prepare_data(model, make_linear(1000))
    rng = np.random.default_rng(21)
    flat = model.initialparameters(rng)
    perm = np.concatenate([rng.permutation(1000) for _ in range(5)])
    o = orc.Oracle(model, opt=eh.Adam(0.001))
    want, ref = _truth_trajectory(o, flat, xf, y, perm, 100, 1)
    sess = eh.FusedSession(model, opt=eh.Adam(0.001))
    sess.upload(0, xf, y)
    sess.set_params(flat)
    got = sess.epoch(perm, 100)
    rel = np.abs(got - want) / np.abs(want)
    assert got.shape == (50,) and rel.max() <= 1e-5, (int(rel.argmax()), float(rel.max()))
    # the global parameter b (the last flat entry, sigmoid-squashed into [0, 10]) after the 50 steps
    b = lambda p: 10.0 / (1.0 + np.exp(-float(p[-1])))
    assert abs(b(sess.get_params()) - b(ref)) <= 1e-4 * b(ref)
    sess.close()


def _blocks(model):
    out, off = [], 0
    for li, (o, i) in enumerate(model.layer_shapes()[0]):
        out.append((f"W{li + 1}", slice(off, off + o * i)))
        off += o * i
        out.append((f"b{li + 1}", slice(off, off + o)))
        off += o
    out.append(("phi", slice(off, model.num_params())))
    return out


def _c5_model(eh):
    return eh.constructHybridModel({"Resp0": ["SM"]}, ["T"], ["Resp_obs", "Resp_obs2"], eh.Expo_resp_model2,
                                   dict(k=(0.01, 0.0, 0.2), Resp0=(2.0, 0.0, 8.0)), ["k"],
                                   hidden_layers=[512, 512, 512], activation="tanh", scale_nn_outputs=False)


def test_c5_wide_batch_65536_loss_gradient_and_10_steps(eh, orc):
    """BASELINE config 5 at its own batch size: [512,512,512], PerTarget(nseLoss, mse), batch 65 536 -- loss and
    gradient of one batch (element-wise bound per parameter block), then 10 Adam steps against the oracle"""
    model = _c5_model(eh)
    n = 1 << 18
    xf, y = eh.prepare_data(model, make_expo2(n))
    rng = np.random.default_rng(31)
    flat = model.initialparameters(rng)
    loss = eh.PerTarget("nseLoss", "mse")
    nt = orc.max_threads()
    o = orc.Oracle(model, training_loss=loss, agg="sum", opt=eh.Adam(0.001))
    sess = eh.FusedSession(model, training_loss=loss, agg="sum", opt=eh.Adam(0.001))
    assert sess.kernel_variant().startswith("wide/")
    sess.upload(0, xf, y)
    sess.set_params(flat)
    idx = rng.permutation(n)[:B_C3]
    L, g = sess.loss_grad(idx)
    L64, g64 = o.loss_grad(flat, xf, y, idx, precision=64, nthreads=nt)
    assert abs(L - L64) <= 2e-2 * abs(L64), (L, L64)
    worst = {}
    for name, sl in _blocks(model):
        a, b = g[sl].astype(np.float64), g64[sl]
        scale = np.abs(b).max()
        if scale < 1e-12:
            continue
        worst[name] = float(np.abs(a - b).max() / scale)
        cos = float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))
        assert cos >= 0.995, (name, cos)
    assert max(worst.values()) <= 4e-2, worst
    # 10 optimiser steps, Float32 oracle trajectory (5 s of CPU per step on 8 cores)
    perm = rng.permutation(n)[: 3 * B_C3]
    perm = np.concatenate([perm, rng.permutation(n)[: 3 * B_C3], rng.permutation(n)[: 3 * B_C3], rng.permutation(n)[:B_C3]])
    got = sess.epoch(perm, B_C3)
    ref = flat.copy()
    want = o.train_steps(ref, xf, y, perm, B_C3, nthreads=nt)
    assert got.shape == (10,)
    rel = np.abs(got - want) / np.abs(want)
    assert rel.max() <= 2e-2, (rel, got, want)
    kphi = lambda p: 0.2 / (1.0 + np.exp(-float(p[-1])))
    assert abs(kphi(sess.get_params()) - kphi(ref)) <= 2e-2 * kphi(ref)
    sess.close()
