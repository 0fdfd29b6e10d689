"""Parity of the CUDA path (through the C ABI) against the oracle on the same seeded inputs.

Tolerances (from BASELINE.json north_star): per-step loss and gradient within 1e-5 relative
(gradient: max-norm relative), trained phi (Q10) within 1e-4 relative after a fixed number of steps.
The float64 oracle is the truth; the float32 oracle shows what the reference's own Float32 path
can achieve on the same inputs."""
import numpy as np
import pytest

from conftest import expo_model, linear_model, make_expo, make_linear, make_synth, rbq10_model, truth_trajectory

pytestmark = pytest.mark.gpu

RTOL_LOSS = 1e-5
RTOL_GRAD = 1e-5
# Single-sample batches: the seed 2 (yhat - y) c of ONE sample carries the whole Float32 rounding of yhat (no averaging),
# so the Float32 reference itself misses 1e-5 there (measured: rbq10-swish, B = 1: Float32 oracle 1.82e-5, GPU 1.82e-5 --
# gpurun_out / profiles/r2_parity_table.txt).  Every batch of two or more samples is held to 1e-5 without exception.
RTOL_GRAD_SINGLE_SAMPLE = 2.5e-5

CASES = [
    ("rbq10-tanh", lambda eh: rbq10_model(eh), lambda: make_synth(4000), "mse", "sum"),
    ("rbq10-tanh-nan", lambda eh: rbq10_model(eh), lambda: make_synth(3000, nan_frac=0.05), "mse", "sum"),
    ("rbq10-swish", lambda eh: rbq10_model(eh, activation="swish"), lambda: make_synth(2000), "mse", "sum"),
    ("rbq10-sigmoid-rmse", lambda eh: rbq10_model(eh, activation="sigmoid"), lambda: make_synth(2000), "rmse", "sum"),
    ("rbq10-noscale-mae", lambda eh: rbq10_model(eh, scale=False), lambda: make_synth(2000), "mae", "sum"),
    ("rbq10-32", lambda eh: rbq10_model(eh, hidden=(32, 32)), lambda: make_synth(2000), "mse", "sum"),
    ("rbq10-bn", lambda eh: rbq10_model(eh, bn=True), lambda: make_synth(2000), "mse", "sum"),
    ("expo-nse", lambda eh: expo_model(eh), lambda: make_expo(500), "nseLoss", "sum"),
    ("expo-bn-nse", lambda eh: expo_model(eh, bn=True), lambda: make_expo(500), "nseLoss", "sum"),
    ("linear-relu", lambda eh: linear_model(eh), lambda: make_linear(1000), "mse", "sum"),
    ("linear2-mean", lambda eh: linear_model(eh, two=True, activation="tanh"), lambda: make_linear(1000, two=True), "mse", "mean"),
    ("linear2-pertarget", lambda eh: linear_model(eh, two=True, activation="tanh"), lambda: make_linear(1000, two=True), "PT", "sum"),
]


def _setup(eh, orc, mk, mkdata, loss, agg, opt=None, seed=7, flags=0):
    model = mk(eh)
    if loss == "PT":
        loss = eh.PerTarget("nseLoss", "mse")
    xf, y = eh.prepare_data(model, mkdata())
    rng = np.random.default_rng(seed)
    flat = model.initialparameters(rng)
    flat += (0.05 * rng.standard_normal(flat.size)).astype(np.float32)
    sess = eh.FusedSession(model, training_loss=loss, agg=agg, opt=opt, flags=flags)
    sess.upload(0, xf, y)
    sess.set_params(flat)
    o = orc.Oracle(model, training_loss=loss, agg=agg, opt=opt)
    return model, xf, y, flat, sess, o, rng


# engine: 0 = default exact-fp32 FFMA2 engine, 16 = EH_FLAG_TENSOR_PIPE (HMMA 3xTF32 engine where the shape has one)
@pytest.mark.parametrize("engine_flags", [0, 16], ids=["ffma2", "mma3xtf32"])
@pytest.mark.parametrize("name,mk,mkdata,loss,agg", CASES, ids=[c[0] for c in CASES])
def test_loss_and_gradient(eh, orc, name, mk, mkdata, loss, agg, engine_flags):
    model, xf, y, flat, sess, o, rng = _setup(eh, orc, mk, mkdata, loss, agg, flags=engine_flags)
    n = xf[0].shape[0]
    for B in (n, 517, 64, 12, 1):  # full, ragged, one chunk, the reference's test batch size, single sample
        if B == 1 and loss in ("nseLoss", "PT"):
            continue  # SS_tot of one sample is 0: the reference yields NaN/Inf there as well
        idx = rng.permutation(n)[:B]
        if np.isnan(np.stack([y[t][idx] for t in model.targets])).all():
            continue
        L, g = sess.loss_grad(idx)
        L64, g64 = o.loss_grad(flat, xf, y, idx, precision=64)
        scale = np.abs(g64).max()
        assert abs(L - L64) <= RTOL_LOSS * abs(L64), (name, B, L, L64)
        err = np.abs(g - g64).max() / scale
        assert err <= (RTOL_GRAD_SINGLE_SAMPLE if B == 1 else RTOL_GRAD), (name, B, err)
    sess.close()


def test_training_trajectory_rbq10(eh, orc):
    """50 Adam steps, batch 512 (README.md:200), per-step loss and final Q10"""
    model, xf, y, flat, sess, o, rng = _setup(eh, orc, lambda e: rbq10_model(e), lambda: make_synth(20000), "mse", "sum",
                                              opt=None)
    perm = np.concatenate([rng.permutation(20000), rng.permutation(20000)])[: 50 * 512]
    got = sess.epoch(perm, 512)
    # EVERY step within 1e-5 of the float64-gradient trajectory (the Float32 oracle, i.e. the reference's own precision,
    # stays within 2e-6 of it on this run)
    want, ref = truth_trajectory(o, flat, xf, y, perm, 512)
    rel = np.abs(got - want) / np.abs(want)
    assert rel.max() <= 1e-5, (int(rel.argmax()), float(rel.max()))
    ps = sess.get_params()
    q10 = lambda p: 1.0 + 3.0 / (1.0 + np.exp(-float(p[-1])))
    assert abs(q10(ps) - q10(ref)) <= 1e-4 * q10(ref)
    # Adam turns a gradient entry at rounding-noise level into a +-eta step, so raw weights are only
    # compared where the gradient is well above noise; the trained model is compared through its output
    _, g0 = o.loss_grad(flat, xf, y, perm[:512], precision=64)
    robust = np.abs(g0) > 1e-2 * np.abs(g0).max()
    assert np.abs(ps - ref)[robust].max() <= 2e-4 * max(1.0, np.abs(ref).max())
    sess.upload(1, xf, y)
    yh = sess.eval(1)[0]
    np.testing.assert_allclose(yh, o.forward(ref, xf, precision=32), rtol=2e-3, atol=2e-3)
    m, v, t = sess.get_opt_state()
    assert t == 50 and np.allclose(m, o.m, atol=1e-4 * np.abs(o.m).max()) and np.allclose(v, o.v, atol=1e-4 * np.abs(o.v).max())
    sess.close()


@pytest.mark.parametrize("optname", ["AdamW", "RMSProp", "Descent"])
def test_other_optimisers(eh, orc, optname):
    opt = {"AdamW": eh.AdamW(0.01, (0.9, 0.999), 0.01), "RMSProp": eh.RMSProp(0.001), "Descent": eh.Descent(0.001)}[optname]
    model, xf, y, flat, sess, o, rng = _setup(eh, orc, lambda e: expo_model(e), lambda: make_expo(500), "nseLoss", "sum", opt=opt)
    perm = rng.permutation(500)
    got = sess.epoch(perm, 64)          # 8 steps, last one partial (52 samples)
    ref = flat.copy()
    want = o.train_steps(ref, xf, y, perm, 64)
    assert got.shape == (8,)
    np.testing.assert_allclose(got, want, rtol=1e-4)
    tol = 1e-4 if optname == "Descent" else 3e-2   # sign-like rules amplify noise-level gradient entries
    np.testing.assert_allclose(sess.get_params(), ref, atol=tol * max(1.0, np.abs(ref).max()))
    assert abs(float(sess.get_params()[-1]) - float(ref[-1])) <= 1e-4 * max(1.0, abs(float(ref[-1])))
    sess.close()


def test_all_masked_batch_is_skipped(eh, orc):
    """src/training/epoch.jl:17-19"""
    table = make_synth(256)
    table["reco"][:64] = np.nan
    model = rbq10_model(eh)
    xf, y = (np.stack([table["sw_pot"], table["dsw_pot"]], 1), {"ta": table["ta"]}), {"reco": table["reco"]}
    flat = model.initialparameters(np.random.default_rng(2))
    sess = eh.FusedSession(model)
    sess.upload(0, xf, y)
    sess.set_params(flat)
    o = orc.Oracle(model)
    got = sess.epoch(np.arange(256), 64)
    ref = flat.copy()
    want = o.train_steps(ref, xf, y, np.arange(256), 64)
    assert np.isnan(got[0]) and np.isnan(want[0])
    np.testing.assert_allclose(got[1:], want[1:], rtol=1e-4)
    assert sess.get_opt_state()[2] == 3
    assert abs(float(sess.get_params()[-1]) - float(ref[-1])) <= 1e-4
    sess.close()


def test_step_host_equals_step_on_indices(eh, orc):
    """collect_dim_data |> gdev path == resident-dataset path, bit for bit"""
    model, xf, y, flat, sess, o, rng = _setup(eh, orc, lambda e: rbq10_model(e), lambda: make_synth(3000, nan_frac=0.03), "mse", "sum")
    idx = rng.permutation(xf[0].shape[0])[:700]
    L1, g1 = sess.step(idx, want_grad=True)
    p1 = sess.get_params()
    sess.set_params(flat)
    sess.set_opt_state(None, None, 0)
    xb = (xf[0][idx], {"ta": xf[1]["ta"][idx]})
    L2 = sess.step_host(xb, {"reco": y["reco"][idx]})
    p2 = sess.get_params()
    assert L1 == L2 and np.array_equal(p1, p2)
    sess.close()


def test_eval_matches_oracle_forward_and_metrics(eh, orc):
    model, xf, y, flat, sess, o, rng = _setup(eh, orc, lambda e: rbq10_model(e), lambda: make_synth(5001, nan_frac=0.1), "mse", "sum")
    sess.upload(1, xf, y)
    yhat, stats, par = sess.eval(1, want_yhat=True, want_params=True)
    want, wpar = o.forward(flat, xf, precision=64, want_params=True)
    np.testing.assert_allclose(yhat, want, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(par[0], wpar[0], rtol=1e-5)          # rb, the NEURAL parameter
    mask = ~np.isnan(y["reco"])
    got = eh.metrics_from_stats(stats[0])
    for kind in eh.LOSS_TYPES:
        assert got[kind] == pytest.approx(orc.loss_fn(want[0], y["reco"], mask, kind), rel=2e-5, abs=1e-7), kind
    sess.close()


def test_determinism_and_additivity_at_benchmark_batch(eh):
    """size-independent properties at the BASELINE batch size (65536): the same step twice is
    bit-identical (atomic-free fixed-order reduction), and n*grad is additive over a split batch"""
    model = rbq10_model(eh)
    t = make_synth(1 << 18)
    xf, y = eh.prepare_data(model, t)
    flat = model.initialparameters(np.random.default_rng(0))
    sess = eh.FusedSession(model)
    sess.upload(0, xf, y)
    sess.set_params(flat)
    idx = np.random.default_rng(1).permutation(1 << 18)[:65536]
    L, g = sess.loss_grad(idx)
    L2, g2 = sess.loss_grad(idx)
    assert L == L2 and np.array_equal(g, g2)
    La, ga = sess.loss_grad(idx[:30000])
    Lb, gb = sess.loss_grad(idx[30000:])
    comb = (30000 * ga.astype(np.float64) + 35536 * gb.astype(np.float64)) / 65536
    assert np.abs(comb - g).max() <= 2e-6 * np.abs(g).max()
    assert abs((30000 * La + 35536 * Lb) / 65536 - L) <= 2e-6 * L
    sess.close()


def test_persistent_kernel_matches_step_kernels(eh, orc):
    """the persistent multi-step kernel (default) and the one-launch-per-step path (EH_FLAG_NO_PERSIST),
    with and without CUDA graph / PDL, train the same model: same losses, same parameters"""
    model = rbq10_model(eh)
    xf, y = eh.prepare_data(model, make_synth(30000, nan_frac=0.02))
    n = xf[0].shape[0]
    flat = model.initialparameters(np.random.default_rng(3))
    perm = np.random.default_rng(4).permutation(n)
    out = {}
    for flags in (0, 4, 4 | 1, 4 | 1 | 2, 16, 16 | 4):
        sess = eh.FusedSession(model, flags=flags)
        sess.upload(0, xf, y)
        sess.set_params(flat)
        sess.set_perm(perm)
        l1 = sess.run_steps(1000, 0, 2 * ((n + 999) // 1000) + 3)   # two passes (graph replay) + 3 steps
        out[flags] = (l1, sess.get_params(), sess.get_opt_state())
        sess.close()
    o = orc.Oracle(model)
    ref = flat.copy()
    nb = (n + 999) // 1000
    want = np.concatenate([o.train_steps(ref, xf, y, perm, 1000), o.train_steps(ref, xf, y, perm, 1000),
                           o.train_steps(ref, xf, y, perm[:3000], 1000)])
    for flags, (l1, ps, (m, v, t)) in out.items():
        assert t == 2 * nb + 3, flags
        np.testing.assert_allclose(l1, want, rtol=5e-4, err_msg=str(flags))
        assert abs(float(ps[-1]) - float(ref[-1])) <= 1e-4, flags
    # the three non-persistent variants share the reduction order: bitwise equal
    assert np.array_equal(out[4][0], out[5][0]) and np.array_equal(out[4][1], out[7][1])


@pytest.mark.parametrize("B,nb", [(8192, 19), (1000, 19), (2048, 53)])
def test_step_host_async_stream_equals_epoch(eh, orc, B, nb):
    # (default: ramped grouped launches.  The opt-in consumer mode, EH_HOST_STREAM=1, takes the same test when the
    # variable is set in the environment of the test run; every host buffer is page-locked BEFORE the burst opens, because
    # a device-synchronising call inside an open burst would deadlock against the waiting kernel)
    """streaming host batches (collect_dim_data |> gdev per step, src/training/epoch.jl:1-11) == the resident path
    on the same batches: per-step losses and trained parameters; NaN targets exercise the per-batch counts.
    53 batches: more than three groups of 16, i.e. the staging ring wraps around"""
    model, xf, y, flat, sess, o, rng = _setup(eh, orc, lambda e: rbq10_model(e), lambda: make_synth(nb * B + 77, nan_frac=0.03), "mse", "sum")
    n = xf[0].shape[0]   # (rows whose only target is NaN were dropped by prepare_data: the last batch is ragged)
    nb = (min(n, nb * B) + B - 1) // B
    perm = rng.permutation(n)[: nb * B]
    want = sess.epoch(perm, B)
    p_want = sess.get_params()
    sess.set_params(flat)
    sess.set_opt_state(None, None, 0)
    losses = sess.pinned(np.zeros(nb, dtype=np.float32))
    keep = []
    for k in range(nb):
        idx = perm[k * B:(k + 1) * B]
        keep.append(sess.host_batch(sess.pinned(xf[0][idx]), [sess.pinned(xf[1]["ta"][idx])], [sess.pinned(y["reco"][idx])]))
    for k in range(nb):
        sess.step_host_async(keep[k], losses, k)
    sess.sync()
    np.testing.assert_allclose(np.asarray(losses), want, rtol=2e-6)
    np.testing.assert_allclose(sess.get_params(), p_want, rtol=0, atol=2e-6)
    assert sess.get_opt_state()[2] == nb
    sess.close()


HOST_CASES = [
    ("rbq10-nan", lambda eh: rbq10_model(eh), lambda n: make_synth(n, nan_frac=0.04), "mse"),
    ("rbq10-bn", lambda eh: rbq10_model(eh, bn=True), lambda n: make_synth(n, nan_frac=0.02), "mse"),
    ("expo-nse", lambda eh: expo_model(eh), lambda n: make_expo(n), "nseLoss"),
    ("linear2", lambda eh: linear_model(eh, two=True, activation="tanh"), lambda n: make_linear(n, two=True), "mse"),
]


@pytest.mark.parametrize("name,mk,mkdata,loss", HOST_CASES, ids=[c[0] for c in HOST_CASES])
def test_zero_copy_host_batches_equal_copy_engine_path(eh, orc, monkeypatch, name, mk, mkdata, loss):
    """page-locked host batches are packed in place over PCIe (k_pack_host); pageable ones -- or all of them with
    EH_HOST_NO_ZEROCOPY=1 -- are staged by the copy engine.  Same records, same per-batch scalars: bit-identical
    losses and parameters, also with NaN targets (valid counts), BatchNorm / nseLoss (K0 on the packed records),
    records wider than one float4, ragged batch sizes."""
    sizes = [4096, 1000, 33, 2048, 5]
    n = sum(sizes)
    results = []
    # 0: default (zero copy; batches without per-batch data statistics are grouped into persistent launches),
    # 1: zero copy, one step / update launch pair per batch, 2: copy engine, one launch pair per batch
    for nogroups, nozc in (("0", "0"), ("1", "0"), ("1", "1")):
        monkeypatch.setenv("EH_HOST_NO_GROUPS", nogroups)
        monkeypatch.setenv("EH_HOST_NO_ZEROCOPY", nozc)
        # (prepare_data drops rows without any valid target: generate more than needed)
        model, xf, y, flat, sess, o, rng = _setup(eh, orc, mk, lambda: mkdata(n + n // 4), loss, "sum")
        assert xf[0].shape[0] >= n
        losses = sess.pinned(np.zeros(len(sizes), dtype=np.float32))
        keep, a = [], 0
        for k, b in enumerate(sizes):
            sl = slice(a, a + b)
            a += b
            hb = sess.host_batch(sess.pinned(xf[0][sl]), [sess.pinned(xf[1][f][sl]) for f in model.forcing],
                                 [sess.pinned(y[t][sl]) for t in model.targets])
            keep.append(hb)
            sess.step_host_async(hb, losses, k)
        sess.sync()
        results.append((np.array(losses), sess.get_params()))
        sess.close()
    assert np.array_equal(results[1][0], results[2][0])
    assert np.array_equal(results[1][1], results[2][1])
    assert np.isfinite(results[0][0]).all()
    # the persistent kernel sums in another order than the step / update pair: same losses to rounding
    np.testing.assert_allclose(results[0][0], results[2][0], rtol=1e-5)


def test_unsupported_models_fail_loudly(eh):
    from easyhybrid_b200 import _abi

    def other(*, ta, Q10, rb):
        return {"reco": rb * np.exp(Q10) + ta}
    # chains wider than 32 run on the tensor-core path, which has no swish
    m = eh.constructHybridModel(["sw_pot"], ["ta"], ["reco"], other, dict(Q10=(2, 1, 4), rb=(3, 0, 13)), ["rb"], ["Q10"],
                                hidden_layers=[64, 64], activation="swish")
    with pytest.raises(eh.EasyHybridCudaError) as ei:
        eh.FusedSession(m)
    assert ei.value.status == _abi.EH_EUNSUPPORTED
    # the same traced model on a narrow chain: generic exact-fp32 variant, any activation
    m = eh.constructHybridModel(["sw_pot"], ["ta"], ["reco"], other, dict(Q10=(2, 1, 4), rb=(3, 0, 13)), ["rb"], ["Q10"],
                                activation="swish")
    sess = eh.FusedSession(m)
    assert sess.kernel_variant().startswith("ffma2/PmProgram/"), sess.kernel_variant()
    sess.close()
    m = eh.constructHybridModel(["sw_pot"], ["ta"], ["reco"], other, dict(Q10=(2, 1, 4), rb=(3, 0, 13)), ["rb"], ["Q10"],
                                hidden_layers=[64, 64])
    sess = eh.FusedSession(m)
    assert sess.kernel_variant().startswith("wide/"), sess.kernel_variant()
    sess.close()
    with pytest.raises(eh.EasyHybridCudaError) as ei:
        eh.FusedSession(rbq10_model(eh, hidden=(600, 600)))   # wider than the tensor-core path takes
    assert ei.value.status == _abi.EH_EUNSUPPORTED
    # everything between the register-tile variants and 512 is served by the tensor-core (bf16 tcgen05) path
    eh.FusedSession(rbq10_model(eh, hidden=(512, 512, 512))).close()
    eh.FusedSession(rbq10_model(eh, hidden=(64, 48))).close()


def test_train_api_learns_q10(eh):
    """train(model, data; ...) end to end: Q10 recovered from the synthetic table (true value 2)"""
    model = rbq10_model(eh, activation="tanh", bn=True)
    out = eh.train(model, make_synth(8192), nepochs=30, batchsize=256, opt=eh.Adam(0.01), loss_types=["mse", "r2", "nse"])
    assert out is not None and len(out.train_history) == 31
    assert out.val_history[-1]["mse"]["sum"] < out.val_history[0]["mse"]["sum"] * 0.05
    assert abs(out.train_diffs["Q10"] - 2.0) < 0.15
    assert out.ps.shape == (model.num_params(),) and out.best_epoch > 0


@pytest.mark.parametrize("B,nan_frac,bn,act", [(65536, 0.0, False, "tanh"), (16384 + 77, 0.03, False, "tanh"), (40000, 0.02, True, "sigmoid")],
                         ids=["c3-batch", "ragged-nan", "bn-sigmoid"])
def test_tensor_engine_gradient_and_trajectory(eh, orc, monkeypatch, B, nan_frac, bn, act):
    """Engine 4 (tcgen05: hidden-layer products with TMEM operands, csrc/eh_engine_tc.cuh) serves the persistent kernel from
    the batch size EH_TC_MIN_BATCH on (opt-in).  (a) Its GRADIENT against the float64 oracle: one Descent(eta) step moves theta by eta * g, so
    (theta0 - theta1) / eta is the engine's gradient, compared at 1e-5 of the max-norm like every other engine (eta is a
    power of two and theta0 is rounded so that the subtraction is exact to ~2^-24 of |theta|: the bound below adds that).
    (b) 12 Adam steps: per-step losses within 1e-5 of the FFMA2 engine (EH_NO_TC=1) and of each other's parameters
    (single_train_step!, src/training/epoch.jl:20-26)."""
    monkeypatch.setenv("EH_TC_MIN_BATCH", "16384")
    model = rbq10_model(eh, bn=bn, activation=act)
    n = 12 * B
    xf, y = eh.prepare_data(model, make_synth(n + 500, nan_frac=nan_frac))
    n = xf[0].shape[0]
    rng = np.random.default_rng(5)
    flat = model.initialparameters(rng)
    flat += (0.05 * rng.standard_normal(flat.size)).astype(np.float32)
    perm = rng.permutation(n)[: min(n, 12 * B)]
    # (a) gradient through one plain gradient-descent step
    eta = 2.0 ** -6
    sess = eh.FusedSession(model, opt=eh.Descent(eta))
    assert "tcgen05" in sess.epoch_variant(B), sess.epoch_variant(B)
    sess.upload(0, xf, y)
    sess.set_params(flat)
    L = sess.epoch(perm[:B], B)
    g = (flat.astype(np.float64) - sess.get_params().astype(np.float64)) / eta
    sess.close()
    o = orc.Oracle(model, opt=eh.Descent(eta))
    L64, g64 = o.loss_grad(flat, xf, y, perm[:B], precision=64, nthreads=orc.max_threads())
    # phi entries: the optimiser moves the raw (logit) parameters, the oracle's gradient is with respect to them as well
    assert abs(L[0] - L64) <= 1e-5 * abs(L64), (L, L64)
    resolution = np.abs(flat).max() * 2.0 ** -23 / eta        # what one float32 update step can resolve
    err = np.abs(g - g64).max()
    assert err <= 1e-5 * np.abs(g64).max() + resolution, (err, np.abs(g64).max(), resolution)
    # (b) trajectory against the FFMA2 engine
    out = []
    for no_tc in (False, True):
        if no_tc:
            monkeypatch.setenv("EH_NO_TC", "1")
        sess = eh.FusedSession(model, opt=eh.Adam(0.01))
        assert ("tcgen05" in sess.epoch_variant(B)) == (not no_tc)
        sess.upload(0, xf, y)
        sess.set_params(flat)
        out.append((sess.epoch(perm, B), sess.get_params()))
        sess.close()
    np.testing.assert_allclose(out[0][0], out[1][0], rtol=1e-5)
    # Adam turns noise-level gradient entries into +-eta steps: compare phi and the bulk of theta
    assert abs(float(out[0][1][-1]) - float(out[1][1][-1])) <= 1e-4
    assert np.median(np.abs(out[0][1] - out[1][1])) <= 1e-4


STAT_CASES = [
    ("rbq10-pearson-nan", lambda eh: rbq10_model(eh), lambda: make_synth(3000, nan_frac=0.05), "pearsonLoss", "sum"),
    ("rbq10-kge", lambda eh: rbq10_model(eh, activation="sigmoid"), lambda: make_synth(3000), "kgeLoss", "sum"),
    ("expo-pbkge-bn", lambda eh: expo_model(eh, bn=True), lambda: make_expo(800), "pbkgeLoss", "sum"),
    ("linear2-rmse-two-targets", lambda eh: linear_model(eh, two=True, activation="tanh"), lambda: make_linear(1500, two=True), "rmse", "mean"),
    ("linear2-kge-mse", lambda eh: linear_model(eh, two=True, activation="tanh"), lambda: make_linear(1500, two=True), "PT-kge-mse", "sum"),
]


@pytest.mark.parametrize("name,mk,mkdata,loss,agg", STAT_CASES, ids=[c[0] for c in STAT_CASES])
def test_prediction_statistics_losses(eh, orc, name, mk, mkdata, loss, agg):
    """rmse over several targets, pearsonLoss, kgeLoss, pbkgeLoss (src/losses/loss_fn.jl:58-60, 75-77, 104-127, 160-174): their
    seeds depend on the mean / variance / covariance of the batch's predictions, so every step runs a forward pre-pass
    (k_eval over the batch + k_stat_seeds) and then the ordinary step kernels with affine seeds.  Loss and gradient at
    1e-5 of the float64 checker; a 12-step Adam trajectory; the host-batch form walks the same trajectory."""
    model = mk(eh)
    tl = eh.PerTarget("kgeLoss", "mse") if loss == "PT-kge-mse" else loss
    xf, y = eh.prepare_data(model, mkdata())
    n = xf[0].shape[0]
    rng = np.random.default_rng(11)
    flat = model.initialparameters(rng)
    flat += (0.05 * rng.standard_normal(flat.size)).astype(np.float32)
    sess = eh.FusedSession(model, training_loss=tl, agg=agg)
    sess.upload(0, xf, y)
    sess.set_params(flat)
    o = orc.Oracle(model, training_loss=tl, agg=agg)
    for B in (n, 517, 64):
        idx = rng.permutation(n)[:B]
        L, g = sess.loss_grad(idx)
        L64, g64 = o.loss_grad(flat, xf, y, idx, precision=64)
        assert abs(L - L64) <= RTOL_LOSS * abs(L64), (name, B, L, L64)
        assert np.abs(g - g64).max() <= RTOL_GRAD * np.abs(g64).max(), (name, B, np.abs(g - g64).max() / np.abs(g64).max())
    B = 256
    perm = rng.permutation(n)[: 12 * B]
    got = sess.epoch(perm, B)
    p_epoch = sess.get_params()
    want, ref = truth_trajectory(o, flat, xf, y, perm, B)
    np.testing.assert_allclose(got, want, rtol=2e-4)
    # the same batches handed over as host arrays (collect_dim_data |> gdev per step)
    sess.set_params(flat)
    sess.set_opt_state(None, None, 0)
    host = []
    for k in range((perm.size + B - 1) // B):   # (the smallest data set has fewer than 12 batches; its last one is ragged)
        idx = perm[k * B:(k + 1) * B]
        host.append(sess.step_host((xf[0][idx], {f: xf[1][f][idx] for f in model.forcing}), {t: y[t][idx] for t in model.targets}))
    np.testing.assert_allclose(host, got, rtol=2e-6)
    np.testing.assert_allclose(sess.get_params(), p_epoch, rtol=0, atol=2e-6)
    sess.close()


@pytest.mark.parametrize("agg,branches,normalize", [("sum", None, True), ("mean", None, False), ("sum", ["Q10"], True)],
                         ids=["sum-all-normalized", "mean-all", "sum-one-branch"])
def test_weight_l2_extra_loss(eh, orc, agg, branches, normalize):
    """The reference's documented extra loss, extra_loss = (ŷ, ps) -> (; l2 = λ * weight_l2(ps.<branch>; normalize),)
    (src/utils/extract_weights.jl:55-91, hook src/losses/compute_loss.jl:31-34), as a native term of the update kernel:
    loss = agg([L, l2]).  Loss and gradient at 1e-5 of the float64 checker, a 10-step trajectory, and train() takes it."""
    from conftest import rbq10_two_chain_model
    model = rbq10_two_chain_model(eh) if branches else rbq10_model(eh)
    xl = eh.WeightL2(0.3, branches=branches, normalize=normalize)
    xf, y = eh.prepare_data(model, make_synth(3000, nan_frac=0.03))
    n = xf[0].shape[0]
    rng = np.random.default_rng(5)
    flat = model.initialparameters(rng)
    flat += (0.1 * rng.standard_normal(flat.size)).astype(np.float32)
    sess = eh.FusedSession(model, agg=agg, extra_loss=xl)
    sess.upload(0, xf, y)
    sess.set_params(flat)
    o = orc.Oracle(model, agg=agg, extra_loss=xl)
    for B in (n, 517, 64):
        idx = rng.permutation(n)[:B]
        L, g = sess.loss_grad(idx)
        L64, g64 = o.loss_grad(flat, xf, y, idx, precision=64)
        assert abs(L - L64) <= RTOL_LOSS * abs(L64), (B, L, L64)
        assert np.abs(g - g64).max() <= RTOL_GRAD * np.abs(g64).max(), (B, np.abs(g - g64).max() / np.abs(g64).max())
    perm = rng.permutation(n)[: 10 * 256]
    got = sess.epoch(perm, 256)
    want, ref = truth_trajectory(o, flat, xf, y, perm, 256)
    np.testing.assert_allclose(got, want, rtol=1e-4)
    sess.close()
