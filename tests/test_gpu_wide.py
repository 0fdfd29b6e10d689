"""Wide-hidden-layer path (bf16 tcgen05 GEMMs, BASELINE config 5) against the float64 oracle.

Stated bf16 tolerance: activations, deltas and the hidden weight matrices are rounded to bf16
(8 bits of mantissa, fp32 accumulation in tensor memory), so per-step loss is compared at 2e-2 relative
and the gradient through its direction and size: cosine similarity >= 0.995 and norm within 3e-2 of the
oracle's, per parameter block.  Everything outside the GEMMs (layer 1, output layer, process model, loss,
optimiser, phi) runs in fp32."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL_LOSS = 2e-2
COS_MIN = 0.995
RTOL_NORM = 3e-2


def make_expo2(n, seed=2314, nan_frac=0.0):
    """Expo recipe (projects/ExpoHybrid/ExpoHybridEstim.jl:39-47), second target = 2 x first + noise."""
    rng = np.random.default_rng(seed)
    T = rng.random(n) * 40 - 10
    SM = rng.random(n) * 0.8 + 0.1
    resp = 1.1 * np.exp(-8.0 * (SM - 0.6) ** 2) * np.exp(0.07 * T)
    obs = resp + rng.standard_normal(n) * 0.05 * resp.mean()
    obs2 = 2.0 * resp + rng.standard_normal(n) * 0.05 * resp.mean()
    if nan_frac:
        obs2[rng.random(n) < nan_frac] = np.nan
    return {k: v.astype(np.float32) for k, v in dict(T=T, SM=SM, Resp_obs=obs, Resp_obs2=obs2).items()}


def wide_model(eh, hidden=(512, 512, 512), two=True, activation="tanh", scale=False):
    if two:
        return eh.constructHybridModel({"Resp0": ["SM"]}, ["T"], ["Resp_obs", "Resp_obs2"], eh.Expo_resp_model2,
                                       dict(k=(0.01, 0.0, 0.2), Resp0=(2.0, 0.0, 8.0)), ["k"],
                                       hidden_layers=list(hidden), activation=activation, scale_nn_outputs=scale)
    return eh.constructHybridModel({"Resp0": ["SM"]}, ["T"], ["Resp_obs"], eh.Expo_resp_model,
                                   dict(k=(0.01, 0.0, 0.2), Resp0=(2.0, 0.0, 8.0)), ["k"],
                                   hidden_layers=list(hidden), activation=activation, scale_nn_outputs=scale)


def _blocks(model):
    """(name, slice) of every parameter block of the flat vector"""
    out, off = [], 0
    for li, (o, i) in enumerate(model.layer_shapes()[0]):
        out.append((f"W{li + 1}", slice(off, off + o * i)))
        off += o * i
        out.append((f"b{li + 1}", slice(off, off + o)))
        off += o
    out.append(("phi", slice(off, model.num_params())))
    return out


CASES = [
    ("c5-3x512-pertarget", dict(hidden=(512, 512, 512), two=True), "PT", 2048),
    ("2x256-mse", dict(hidden=(256, 256), two=False, activation="sigmoid"), "mse", 1024),
    ("3x512-scale-nan", dict(hidden=(512, 512, 512), two=True, scale=True), "mse", 1536),
    # widths between the register-tile kernels (<= 32) and 256 / 512 are padded inside the tensor-core path
    ("2x64-mse", dict(hidden=(64, 64), two=False), "mse", 1024),
    ("48-96-32-relu", dict(hidden=(48, 96, 32), two=True, activation="relu"), "mse", 1024),
    ("3x16-deeper-than-the-aot-variants", dict(hidden=(16, 16, 16), two=False), "mse", 1024),
    ("300-200-sigmoid-scale", dict(hidden=(300, 200), two=True, activation="sigmoid", scale=True), "PT", 1024),
]


@pytest.mark.parametrize("name,kw,loss,n", CASES, ids=[c[0] for c in CASES])
def test_wide_loss_and_gradient(eh, orc, name, kw, loss, n):
    model = wide_model(eh, **kw)
    if loss == "PT":
        loss = eh.PerTarget("nseLoss", "mse")
    data = make_expo2(n, nan_frac=0.05 if "nan" in name else 0.0)
    xf, y = eh.prepare_data(model, data)
    rng = np.random.default_rng(3)
    flat = model.initialparameters(rng)
    sess = eh.FusedSession(model, training_loss=loss, agg="sum")
    sess.upload(0, xf, y)
    sess.set_params(flat)
    o = orc.Oracle(model, training_loss=loss, agg="sum")
    for B in (n, 512, 128):
        idx = rng.permutation(n)[:B]
        L, g = sess.loss_grad(idx)
        L64, g64 = o.loss_grad(flat, xf, y, idx, precision=64)
        assert abs(L - L64) <= RTOL_LOSS * abs(L64), (name, B, L, L64)
        for bname, sl in _blocks(model):
            a, b = g[sl].astype(np.float64), g64[sl].astype(np.float64)
            nb = np.linalg.norm(b)
            if nb < 1e-12:
                continue
            cos = float(a @ b / (np.linalg.norm(a) * nb))
            assert cos >= COS_MIN, (name, B, bname, cos)
            assert abs(np.linalg.norm(a) - nb) <= RTOL_NORM * nb, (name, B, bname, np.linalg.norm(a), nb)
    sess.close()


def test_wide_training_reduces_loss_and_tracks_oracle(eh, orc):
    """30 Adam steps on the C5 shape: the loss trajectory follows the float64 oracle and k is learned alike"""
    model = wide_model(eh)
    n, B = 4096, 1024
    data = make_expo2(n)
    xf, y = eh.prepare_data(model, data)
    rng = np.random.default_rng(11)
    flat = model.initialparameters(rng)
    loss = eh.PerTarget("nseLoss", "mse")
    sess = eh.FusedSession(model, training_loss=loss, agg="sum", opt=eh.Adam(0.001))
    sess.upload(0, xf, y)
    sess.set_params(flat)
    o = orc.Oracle(model, training_loss=loss, agg="sum", opt=eh.Adam(0.001))
    perm = np.concatenate([rng.permutation(n) for _ in range(8)])[: 30 * B]
    got = sess.epoch(perm, B)
    ref = flat.copy()
    want = o.train_steps(ref, xf, y, perm, B)
    assert got[-1] < 0.6 * got[0]
    assert np.allclose(got, want, rtol=5e-2), (got, want)
    phi_got, phi_want = float(sess.get_params()[-1]), float(ref[-1])
    assert abs(phi_got - phi_want) <= 2e-2 * max(1.0, abs(phi_want)), (phi_got, phi_want)
    # evaluation (test-mode forward over the whole split) agrees with the oracle's predictions
    yhat, stats, _ = sess.eval(0, want_yhat=True)
    yh = o.forward(sess.get_params(), xf, precision=64)
    assert np.allclose(yhat, yh, rtol=3e-2, atol=3e-2)
    assert stats[0, 0] == n  # valid count of target 0
    mse0 = stats[0, 6] / stats[0, 0]
    assert abs(mse0 - np.mean((yhat[0] - y[model.targets[0]]) ** 2)) <= 1e-3 * mse0
    sess.close()


def test_wide_unsupported_shapes_fail_loudly(eh):
    model = wide_model(eh, hidden=(1024, 1024))
    with pytest.raises(eh.EasyHybridCudaError) as ei:
        eh.FusedSession(model, training_loss="mse")
    assert "EH_EUNSUPPORTED" in str(ei.value) or "no fused kernel" in str(ei.value)
    model = wide_model(eh, hidden=(64, 64), activation="swish")
    with pytest.raises(eh.EasyHybridCudaError):
        eh.FusedSession(model, training_loss="mse")


def test_tcgen05_gemm_kernels_against_torch():
    """the three GEMM kinds in isolation (eh_selftest_wide_gemm): K-major forward / backward-data with fused
    epilogues, MN-major split-K weight gradient"""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "wide_gemm_check", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "wide_gemm_check.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.main(False)


def test_traced_process_model_runs_on_the_tensor_core_path(eh, orc, monkeypatch):
    """a process model that is none of the built-in forms (src/models/GenericHybridModel.jl:425 takes any callable):
    traced into a straight-line program by the host, interpreted (value and pullback) in the head kernel.
    (Chains this narrow normally take the exact-fp32 generic variants, tests/test_gpu_generic.py; EH_NO_SMALL_PROGRAM
    sends the model to the tensor-core path, which is what wider chains with a traced model use.)"""
    from conftest import make_synth
    monkeypatch.setenv("EH_NO_SMALL_PROGRAM", "1")

    def custom(*, ta, dsw_pot, rb, Q10, alpha, tref=15.0):
        return {"reco": rb * Q10 ** (0.1 * (ta - tref)) + alpha * np.tanh(0.05 * dsw_pot)}

    model = eh.constructHybridModel(["sw_pot", "dsw_pot"], ["ta", "dsw_pot"], ["reco"], custom,
                                    dict(rb=(3.0, 0.0, 13.0), Q10=(2.0, 1.0, 4.0), alpha=(0.5, -2.0, 2.0)), ["rb"], ["Q10", "alpha"],
                                    hidden_layers=[32, 32], activation="tanh", scale_nn_outputs=True)
    from easyhybrid_b200 import _abi
    assert eh.build_desc(model).desc.process_model == _abi.PM["PROGRAM"]
    n = 2048
    table = make_synth(n, nan_frac=0.0)
    table["reco"] = (table["reco"] + 0.3 * np.tanh(0.05 * table["dsw_pot"])).astype(np.float32)
    xf, y = eh.prepare_data(model, table)
    rng = np.random.default_rng(5)
    flat = model.initialparameters(rng)
    sess = eh.FusedSession(model, training_loss="mse", opt=eh.Adam(0.01))
    assert sess.kernel_variant().startswith("wide/"), sess.kernel_variant()
    sess.upload(0, xf, y)
    sess.set_params(flat)
    o = orc.Oracle(model, training_loss="mse", opt=eh.Adam(0.01))
    idx = rng.permutation(n)[:1024]
    L, g = sess.loss_grad(idx)
    L64, g64 = o.loss_grad(flat, xf, y, idx, precision=64)
    assert abs(L - L64) <= RTOL_LOSS * abs(L64), (L, L64)
    for bname, sl in _blocks(model):
        a, b = g[sl].astype(np.float64), g64[sl].astype(np.float64)
        nb = np.linalg.norm(b)
        if nb < 1e-12:
            continue
        assert float(a @ b / (np.linalg.norm(a) * nb)) >= COS_MIN, bname
        assert abs(np.linalg.norm(a) - nb) <= RTOL_NORM * nb, bname
    perm = np.concatenate([rng.permutation(n) for _ in range(10)])[: 40 * 512]
    got = sess.epoch(perm, 512)
    ref = flat.copy()
    want = o.train_steps(ref, xf, y, perm, 512)
    assert got[-1] < 0.5 * got[0]
    assert np.allclose(got, want, rtol=5e-2, atol=1e-3), (got[-5:], want[-5:])
    ps = sess.get_params()
    assert np.allclose(ps[-2:], ref[-2:], atol=3e-2), (ps[-2:], ref[-2:])   # Q10 and alpha (raw)
    yhat, stats, par = sess.eval(0, want_yhat=True, want_params=True)
    assert np.allclose(yhat, o.forward(ps, xf, precision=64), rtol=3e-2, atol=3e-2)
    sess.close()


def test_wide_ragged_batches_and_host_batches(eh, orc):
    """batch sizes that are no multiple of the 128-row GEMM tile (partial last batch of an epoch,
    src/data/loaders.jl:1-12 with partial = true) are padded and masked inside; host batches
    (collect_dim_data |> gdev) take the same path"""
    model = wide_model(eh, hidden=(256, 256), two=False)
    n, B = 2500, 700                      # 4 batches: 700, 700, 700, 400
    data = make_expo2(n, nan_frac=0.0)
    xf, y = eh.prepare_data(model, data)
    rng = np.random.default_rng(21)
    flat = model.initialparameters(rng)
    sess = eh.FusedSession(model, training_loss="mse", opt=eh.Adam(0.001))
    sess.upload(0, xf, y)
    sess.set_params(flat)
    o = orc.Oracle(model, training_loss="mse", opt=eh.Adam(0.001))
    for Bq in (333, 1, 129):
        idx = rng.permutation(n)[:Bq]
        L, g = sess.loss_grad(idx)
        L64, g64 = o.loss_grad(flat, xf, y, idx, precision=64)
        assert abs(L - L64) <= RTOL_LOSS * abs(L64), (Bq, L, L64)
        cos = float(g @ g64 / (np.linalg.norm(g) * np.linalg.norm(g64)))
        assert cos >= COS_MIN, (Bq, cos)
    perm = rng.permutation(n)
    got = sess.epoch(perm, B)
    ref = flat.copy()
    want = o.train_steps(ref, xf, y, perm, B)
    assert got.shape == (4,) and np.allclose(got, want, rtol=3e-2), (got, want)
    # the same four batches as host arrays
    sess.set_params(flat)
    sess.set_opt_state(None, None, 0)
    losses = sess.pinned(np.zeros(4, dtype=np.float32))
    keep = []
    for k in range(4):
        idx = perm[k * B:(k + 1) * B]
        hb = sess.host_batch(sess.pinned(xf[0][idx]), [sess.pinned(xf[1]["T"][idx])], [sess.pinned(y["Resp_obs"][idx])])
        keep.append(hb)
        sess.step_host_async(hb, losses, k)
    sess.sync()
    assert np.allclose(np.asarray(losses), got, rtol=1e-5), (np.asarray(losses), got)
    sess.close()


def test_two_chain_model_embedded_block_diagonally(eh, orc):
    """MultiNNHybridModel with a network per parameter (src/models/GenericHybridModel.jl:169-189, forward :458-530): the
    two chains train as one block-diagonally embedded chain on the tensor-core path"""
    from conftest import make_synth, rbq10_two_chain_model
    model = rbq10_two_chain_model(eh, hidden=(24, 16))
    n = 3000
    xf, y = eh.prepare_data(model, make_synth(n, nan_frac=0.04), drop_missing_rows=False)
    rng = np.random.default_rng(9)
    flat = model.initialparameters(rng)
    flat += (0.05 * rng.standard_normal(flat.size)).astype(np.float32)
    sess = eh.FusedSession(model, training_loss="mse", opt=eh.Adam(0.01))
    sess.upload(0, xf, y)
    sess.set_params(flat)
    o = orc.Oracle(model, training_loss="mse", opt=eh.Adam(0.01))
    for B in (n, 500):
        idx = rng.permutation(n)[:B]
        L, g = sess.loss_grad(idx)
        L64, g64 = o.loss_grad(flat, xf, y, idx, precision=64)
        assert abs(L - L64) <= RTOL_LOSS * abs(L64), (B, L, L64)
        # per chain: the gradient blocks of the flat vector
        half = [slice(0, flat.size // 2 + 16), slice(flat.size // 2 + 16, flat.size)]
        for sl in half + [slice(0, flat.size)]:
            a, b = g[sl].astype(np.float64), g64[sl].astype(np.float64)
            cos = float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))
            assert cos >= COS_MIN, (B, sl, cos)
            assert abs(np.linalg.norm(a) - np.linalg.norm(b)) <= RTOL_NORM * np.linalg.norm(b)
    perm = np.concatenate([rng.permutation(n) for _ in range(6)])[: 30 * 512]
    got = sess.epoch(perm, 512)
    ref = flat.copy()
    want = o.train_steps(ref, xf, y, perm, 512)
    assert got[-1] < 0.5 * got[0]
    assert np.allclose(got, want, rtol=5e-2, atol=1e-3), (got[-4:], want[-4:])
    yhat, stats, par = sess.eval(0, want_yhat=True, want_params=True)
    yo, po = o.forward(sess.get_params(), xf, precision=64, want_params=True)
    assert np.allclose(yhat, yo, rtol=3e-2, atol=3e-2)
    assert np.allclose(par, po, rtol=3e-2, atol=3e-2)   # both neural parameters per sample
    sess.close()


def test_six_inputs_over_two_chains(eh, orc):
    """more chain inputs than the four the first version of the path took (both chains see all three columns)"""
    from conftest import make_synth
    model = eh.constructHybridModel({"rb": ["sw_pot", "dsw_pot", "ta"], "Q10": ["ta", "sw_pot", "dsw_pot"]}, ["ta"], ["reco"], eh.RbQ10,
                                    dict(rb=(3.0, 0.0, 13.0), Q10=(2.0, 1.0, 4.0)), [],
                                    hidden_layers=[32, 32], activation="sigmoid", scale_nn_outputs=True, input_batchnorm=True)
    n = 2048
    xf, y = eh.prepare_data(model, make_synth(n))
    rng = np.random.default_rng(13)
    flat = model.initialparameters(rng)
    flat += (0.05 * rng.standard_normal(flat.size)).astype(np.float32)
    sess = eh.FusedSession(model, training_loss="mse")
    sess.upload(0, xf, y)
    sess.set_params(flat)
    o = orc.Oracle(model, training_loss="mse")
    idx = rng.permutation(n)[:1000]
    L, g = sess.loss_grad(idx)
    L64, g64 = o.loss_grad(flat, xf, y, idx, precision=64)
    assert abs(L - L64) <= RTOL_LOSS * abs(L64), (L, L64)
    cos = float(g @ g64 / (np.linalg.norm(g) * np.linalg.norm(g64)))
    assert cos >= COS_MIN, cos
    assert abs(np.linalg.norm(g) - np.linalg.norm(g64)) <= RTOL_NORM * np.linalg.norm(g64)
    sess.close()


def test_train_api_with_a_chain_outside_the_register_tile_variants(eh):
    """train(model, data; ...) (src/training/train.jl) end to end on the tensor-core path: hidden [64, 64], ragged
    batches (batchsize 300), evaluation metrics per epoch, Q10 recovered from the synthetic table (true value 2)"""
    from conftest import make_synth, rbq10_model
    model = rbq10_model(eh, hidden=(64, 64), activation="tanh", bn=True)
    out = eh.train(model, make_synth(6000), nepochs=12, batchsize=300, opt=eh.Adam(0.01), loss_types=["mse", "r2"])
    assert out is not None and len(out.train_history) == 13
    assert out.val_history[-1]["mse"]["sum"] < out.val_history[0]["mse"]["sum"] * 0.1
    assert abs(out.train_diffs["Q10"] - 2.0) < 0.25
