"""Run-time specialisation of traced process models (csrc/eh_jit.cu; SURVEY 8 f3 "traced / NVRTC plug-in").

CPU part: the planner picks the generic exact-fp32 shape, the generated functor compiles with NVRTC for sm_100a and the
result lands in the disk cache -- no device involved (eh_jit_check).  GPU part: the compiled kernels against the CPU
checker at the bar of the interpreted ones (1e-5), against the interpreter itself, and the three launch forms
(single step, persistent epoch, eval)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import test_gpu_generic as gg  # noqa: E402
from conftest import make_synth  # noqa: E402


def _check(eh, model, **kw):
    from easyhybrid_b200._lib import load
    from easyhybrid_b200.model import build_desc
    lib = load()
    bundle = build_desc(model, **kw)
    info = C.create_string_buffer(512)
    st = lib.eh_jit_check(bundle.byref(), info, 512)
    return st, info.value.decode(), (lib.eh_last_error(None) or b"").decode()


def test_traced_program_compiles_without_a_device(eh, tmp_path, monkeypatch):
    monkeypatch.setenv("EH_JIT_CACHE", str(tmp_path))
    st, info, err = _check(eh, gg.m_custom(eh), training_loss="mse", agg="sum", opt=eh.Adam(0.01))
    assert st == 0, err
    assert info.startswith("nvrtc/PmTraced#") and "/P2/NH2/H16/O1/ACT_TANH" in info and " cached=0 " in info, info
    assert int(info.split("cubin=")[1].split()[0]) > 100_000
    assert len([f for f in os.listdir(tmp_path) if f.endswith(".ehjit")]) == 1
    # second time: from the cache, same kernels
    st, info2, err = _check(eh, gg.m_custom(eh), training_loss="mse", agg="sum", opt=eh.Adam(0.01))
    assert st == 0 and " cached=1 " in info2 and info2.split()[0] == info.split()[0], (info2, err)
    # another program / shape -> another entry
    st, info3, err = _check(eh, gg.m_traced_unequal_depth(eh), training_loss="mse", agg="sum")
    assert st == 0 and info3.split()[0] != info.split()[0] and "/NH2/" in info3, (info3, err)


def test_models_off_the_generic_path_are_refused(eh, tmp_path, monkeypatch):
    from conftest import rbq10_model
    monkeypatch.setenv("EH_JIT_CACHE", str(tmp_path))
    st, info, err = _check(eh, rbq10_model(eh), training_loss="mse", agg="sum")   # a built-in form: specialised variant
    assert st != 0 and "did not choose a generic" in err, (st, info, err)


def three_param_pm(*, ta, dsw_pot, rb, Q10, alpha, tref=15.0):
    return {"reco": rb * Q10 ** (0.1 * (ta - tref)) + alpha * np.tanh(0.05 * dsw_pot)}


def m_three_neural(eh):
    # three neural parameters out of ONE chain (SingleNNHybridModel, GenericHybridModel.jl:89-157) of widths 24 / 20: three
    # chain outputs and width 24 exist as run-time compiled shapes only
    return eh.constructHybridModel(["sw_pot", "dsw_pot", "ta"], ["ta", "dsw_pot"], ["reco"], three_param_pm,
                                   dict(rb=(3.0, 0.0, 13.0), Q10=(2.0, 1.0, 4.0), alpha=(0.5, -2.0, 2.0)), ["rb", "Q10", "alpha"], [],
                                   hidden_layers=[24, 20], activation="tanh", scale_nn_outputs=True, input_batchnorm=True)


def _table_wide(n, seed=9):
    t = gg._table(n)
    rng = np.random.default_rng(seed)
    for k in range(7):
        t["z%d" % k] = rng.standard_normal(n).astype(np.float32)
    return t


def m_ten_inputs(eh):
    # ten predictors: inputs padded to 12 (the compiled-in variants stop at 8)
    return eh.constructHybridModel(["sw_pot", "dsw_pot", "ta"] + ["z%d" % k for k in range(7)], ["ta", "dsw_pot"], ["reco"], gg.custom_pm,
                                   dict(rb=(3.0, 0.0, 13.0), Q10=(2.0, 1.0, 4.0), alpha=(0.5, -2.0, 2.0)), ["rb"], ["Q10", "alpha"],
                                   hidden_layers=[16], activation="sigmoid", scale_nn_outputs=True, input_batchnorm=True)


def test_shapes_beyond_the_compiled_in_variants(eh, tmp_path, monkeypatch):
    """three chain outputs / width 24 / ten inputs: EH_EUNSUPPORTED-or-bf16 without run-time compilation, exact fp32 with it"""
    monkeypatch.setenv("EH_JIT_CACHE", str(tmp_path))
    st, info, err = _check(eh, m_three_neural(eh), training_loss="mse", agg="sum")
    assert st == 0 and "/P4/NH2/H24/O3/ACT_TANH" in info, (info, err)
    st, info, err = _check(eh, m_ten_inputs(eh), training_loss="mse", agg="sum")
    assert st == 0 and "/P12/NH1/H16/O1/ACT_SIGMOID" in info, (info, err)
    # the tightest shape is taken where a compiled-in variant would pad: hidden 12 -> width 16 there, 12 -> 16 here too,
    # but 8 -> 8 instead of 16
    st, info, err = _check(eh, gg.m_custom(eh, hidden=(8, 8)), training_loss="mse", agg="sum")
    assert st == 0 and "/P2/NH2/H8/O1/" in info, (info, err)
    # an activation per chain
    st, info, err = _check(eh, gg.m_chains_mixed_activation(eh), training_loss="mse", agg="sum")
    assert st == 0 and info.split()[0].endswith("/P4/NH2/H24/O2/ACT_PER_UNIT"), (info, err)


JIT_CASES = [
    ("traced-all-operations", gg.m_many_ops, lambda: gg._table(2000, nan_frac=0.03), "mse", "sum"),
    ("custom-tanh16", gg.m_custom, lambda: gg._table(3000, nan_frac=0.05), "mse", "sum"),
    ("custom-two-targets-bn", gg.m_two_targets, lambda: gg._table(2500, nan_frac=0.1, two=True), "PT", "mean"),
    ("traced-chains-depth-1-and-2-relu", gg.m_traced_unequal_depth, lambda: gg._table(2000), "mse", "mean"),
    ("three-neural-parameters-width-24", m_three_neural, lambda: gg._table(2000, nan_frac=0.03), "mse", "sum"),
    ("ten-inputs", m_ten_inputs, lambda: _table_wide(2000), "nseLoss", "sum"),
]
JIT_CASES += [
    # chains that differ in activation: per-unit activations in the generated code
    ("chains-tanh-and-relu-depth-2-and-1", gg.m_chains_mixed_activation, lambda: gg._table(2000, nan_frac=0.03), "mse", "sum"),
    ("chains-swish-and-sigmoid", lambda eh: gg.m_chains_mixed_activation(eh, ("swish", "sigmoid"), {"rb": [9, 9], "Q10": [6, 5]}),
     lambda: gg._table(2000), "mae", "sum"),
]
JIT_ONLY = {"three-neural-parameters-width-24", "ten-inputs", "chains-tanh-and-relu-depth-2-and-1", "chains-swish-and-sigmoid"}   # no compiled-in variant: nothing to compare the interpreter with


@pytest.mark.gpu
@pytest.mark.parametrize("name,mk,mkdata,loss,agg", JIT_CASES, ids=[c[0] for c in JIT_CASES])
def test_compiled_program_matches_checker_and_interpreter(eh, orc, name, mk, mkdata, loss, agg):
    model = mk(eh)
    if loss == "PT":
        loss = eh.PerTarget("nseLoss", "mse")
    xf, y = eh.prepare_data(model, mkdata())
    rng = np.random.default_rng(5)
    flat = model.initialparameters(rng)
    flat += (0.05 * rng.standard_normal(flat.size)).astype(np.float32)
    o = orc.Oracle(model, training_loss=loss, agg=agg, opt=eh.Adam(0.01))
    sj = eh.FusedSession(model, training_loss=loss, agg=agg, opt=eh.Adam(0.01), jit=True)
    assert sj.kernel_variant().startswith("nvrtc/PmTraced#"), sj.kernel_variant()
    si = None
    if name not in JIT_ONLY:
        si = eh.FusedSession(model, training_loss=loss, agg=agg, opt=eh.Adam(0.01))
        assert si.kernel_variant().startswith("ffma2/PmProgram/"), si.kernel_variant()
    n = xf[0].shape[0]
    for s in (sj, si):
        if s is None:
            continue
        s.upload(0, xf, y)
        s.set_params(flat)
    for B in (n, 517, 12):
        idx = rng.permutation(n)[:B]
        L, g = sj.loss_grad(idx)
        L64, g64 = o.loss_grad(flat, xf, y, idx, precision=64)
        scale = np.abs(g64).max()
        assert abs(L - L64) <= 2e-6 * abs(L64), (name, B, L, L64)
        assert np.abs(g - g64).max() <= 1e-5 * scale, (name, B)
        if si is not None:
            Li, gi = si.loss_grad(idx)
            assert np.abs(g - gi).max() <= 2e-6 * scale and abs(L - Li) <= 1e-6 * abs(Li)   # same formulas, contraction aside
    # persistent epoch kernel, single step, eval
    B = 256
    perm = np.concatenate([rng.permutation(n) for _ in range(3)])[: 20 * B]
    got = sj.epoch(perm, B)
    ref = flat.copy()
    want = o.train_steps(ref, xf, y, perm, B)
    np.testing.assert_allclose(got, want, rtol=2e-4)
    if si is not None:
        np.testing.assert_allclose(got, si.epoch(perm, B), rtol=2e-5)
    L1 = sj.step(perm[:B])
    Lo = o.train_steps(ref, xf, y, perm[:B], B)
    assert abs(L1 - Lo[0]) <= 2e-4 * abs(Lo[0])
    ps = sj.get_params()
    yhat, stats, par = sj.eval(0, want_yhat=True, want_params=True)
    want_y = o.forward(ps, xf, precision=64)
    assert np.allclose(yhat, want_y, rtol=2e-5, atol=2e-5)
    sj.close()
    if si is not None:
        si.close()


@pytest.mark.gpu
def test_shapes_without_a_compiled_in_variant_compile_unasked(eh, monkeypatch):
    """three chain outputs: no flag needed -- run-time compilation is the only exact-fp32 path; EH_JIT=0 forbids it and the
    model is refused (the tensor-core path takes one or two chain outputs)"""
    sess = eh.FusedSession(m_three_neural(eh))
    assert sess.kernel_variant().startswith("nvrtc/PmTraced#") and "/O3/" in sess.kernel_variant(), sess.kernel_variant()
    sess.close()
    monkeypatch.setenv("EH_JIT", "0")
    with pytest.raises(eh.EasyHybridCudaError):
        eh.FusedSession(m_three_neural(eh))


@pytest.mark.gpu
def test_host_batches_through_compiled_kernels(eh):
    """collect_dim_data |> gdev path (eh_step_host, and the asynchronous grouped form) == resident-dataset path, bit for
    bit, with the run-time compiled kernels"""
    model = gg.m_custom(eh)
    xf, y = eh.prepare_data(model, gg._table(3000, nan_frac=0.03))
    rng = np.random.default_rng(2)
    flat = model.initialparameters(rng)
    sess = eh.FusedSession(model, opt=eh.Adam(0.01), jit=True)
    assert sess.kernel_variant().startswith("nvrtc/")
    sess.upload(0, xf, y)
    sess.set_params(flat)
    B, nb = 500, 6
    perm = rng.permutation(xf[0].shape[0])[: B * nb]
    want = sess.epoch(perm, B)
    p1 = sess.get_params()
    sess.set_params(flat)
    sess.set_opt_state(None, None, 0)
    got = []
    for k in range(nb):
        idx = perm[k * B:(k + 1) * B]
        xb = (xf[0][idx], {f: xf[1][f][idx] for f in model.forcing})
        got.append(sess.step_host(xb, {t: y[t][idx] for t in model.targets}))
    np.testing.assert_allclose(got, want, rtol=2e-6)
    np.testing.assert_allclose(sess.get_params(), p1, rtol=0, atol=2e-6)
    sess.close()


@pytest.mark.gpu
def test_train_api_with_compiled_program(eh):
    table = gg._table(8192)
    res = eh.train(gg.m_custom(eh), table, nepochs=30, batchsize=512, opt=eh.Adam(0.01), training_loss="mse", random_seed=1,
                   patience=100, jit=True)
    assert res.val_history[-1]["mse"]["sum"] < 0.03
    assert abs(res.train_diffs["Q10"] - 2.0) < 0.1
