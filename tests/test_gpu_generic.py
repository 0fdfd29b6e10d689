"""The generic exact-fp32 variants of the register-tile kernels: any process model the host can trace
(src/models/GenericHybridModel.jl:425 takes any callable), and the built-in forms in shapes without a specialised
variant, run on the same fused fp32 kernels as the BASELINE configurations -- the process model is interpreted per
sample (value and reverse sweep), chain inputs are padded with zero columns, unused forcing / target columns of the
compiled maxima are zero / always-masked.  Same tolerances as tests/test_gpu_parity.py: 1e-5 relative against the
float64 oracle."""
import numpy as np
import pytest

from conftest import make_synth, rbq10_model, rbq10_two_chain_model

pytestmark = pytest.mark.gpu

RTOL_LOSS = 1e-5
RTOL_GRAD = 1e-5


def custom_pm(*, ta, dsw_pot, rb, Q10, alpha, tref=15.0):
    return {"reco": rb * Q10 ** (0.1 * (ta - tref)) + alpha * np.tanh(0.05 * dsw_pot)}


def custom_two_targets(*, ta, dsw_pot, rb, Q10, alpha, tref=15.0):
    r = rb * Q10 ** (0.1 * (ta - tref))
    return {"reco": r + alpha * np.tanh(0.05 * dsw_pot), "reco2": 2.0 * r + alpha}


def many_ops_pm(*, ta, dsw_pot, rb, Q10, alpha, tref=15.0):
    """every operation the program format knows except pow / tanh (custom_pm has those): exp, log, sqrt, abs, sin, cos, min,
    max, div, neg, sub -- the reverse sweep of each is exercised (interpreter and run-time compiled code)"""
    t = 0.1 * (ta - tref)
    a = np.exp(t * np.log(Q10))
    b = np.sqrt(rb * rb + 1.0) / (1.0 + np.abs(alpha))
    c = np.sin(0.05 * dsw_pot) * np.cos(alpha)
    d = np.minimum(rb, 4.0) + np.maximum(alpha, -0.25) - (-t)
    return {"reco": a * b + 0.1 * c + 0.01 * d}


def m_many_ops(eh):
    return eh.constructHybridModel(["sw_pot", "dsw_pot"], ["ta", "dsw_pot"], ["reco"], many_ops_pm,
                                   dict(rb=(3.0, 0.0, 13.0), Q10=(2.0, 1.0, 4.0), alpha=(0.5, -2.0, 2.0)), ["rb"], ["Q10", "alpha"],
                                   hidden_layers=[16, 16], activation="tanh", scale_nn_outputs=True, input_batchnorm=True)


def _table(n, nan_frac=0.0, two=False):
    t = make_synth(n, nan_frac=0.0)
    t["reco"] = (t["reco"] + 0.3 * np.tanh(0.05 * t["dsw_pot"])).astype(np.float32)
    if two:
        rng = np.random.default_rng(3)
        t["reco2"] = (2.0 * t["reco"] + 0.3 + 0.05 * rng.standard_normal(n)).astype(np.float32)
        if nan_frac:
            t["reco2"] = np.where(rng.random(n) < nan_frac, np.nan, t["reco2"]).astype(np.float32)
    elif nan_frac:
        rng = np.random.default_rng(4)
        t["reco"] = np.where(rng.random(n) < nan_frac, np.nan, t["reco"]).astype(np.float32)
    return t


def m_custom(eh, hidden=(16, 16), activation="tanh", scale=True, bn=False):
    return eh.constructHybridModel(["sw_pot", "dsw_pot"], ["ta", "dsw_pot"], ["reco"], custom_pm,
                                   dict(rb=(3.0, 0.0, 13.0), Q10=(2.0, 1.0, 4.0), alpha=(0.5, -2.0, 2.0)), ["rb"], ["Q10", "alpha"],
                                   hidden_layers=list(hidden), activation=activation, scale_nn_outputs=scale, input_batchnorm=bn)


def m_two_neural(eh, hidden=(16, 16)):
    return eh.constructHybridModel(["sw_pot", "dsw_pot"], ["ta", "dsw_pot"], ["reco"], custom_pm,
                                   dict(rb=(3.0, 0.0, 13.0), Q10=(2.0, 1.0, 4.0), alpha=(0.5, -2.0, 2.0)), ["rb", "Q10"], ["alpha"],
                                   hidden_layers=list(hidden), activation="sigmoid", scale_nn_outputs=True)


def m_two_targets(eh):
    return eh.constructHybridModel(["sw_pot", "dsw_pot", "ta"], ["ta", "dsw_pot"], ["reco", "reco2"], custom_two_targets,
                                   dict(rb=(3.0, 0.0, 13.0), Q10=(2.0, 1.0, 4.0), alpha=(0.5, -2.0, 2.0)), ["rb"], ["Q10", "alpha"],
                                   hidden_layers=[16, 16], activation="tanh", scale_nn_outputs=True, input_batchnorm=True)


def m_rbq10_three_inputs(eh):
    # a built-in form in a shape without a specialised variant: three predictors, swish
    return eh.constructHybridModel(["sw_pot", "dsw_pot", "ta"], ["ta"], ["reco"], eh.RbQ10,
                                   dict(rb=(3.0, 0.0, 13.0), Q10=(2.0, 1.0, 4.0)), ["rb"], ["Q10"],
                                   hidden_layers=[16, 16], activation="swish", scale_nn_outputs=True)


def m_two_chains_six_inputs(eh):
    # MultiNNHybridModel (GenericHybridModel.jl:169-189): both chains see all three columns, unequal widths, BatchNorm
    return eh.constructHybridModel({"rb": ["sw_pot", "dsw_pot", "ta"], "Q10": ["ta", "sw_pot", "dsw_pot"]}, ["ta"], ["reco"], eh.RbQ10,
                                   dict(rb=(3.0, 0.0, 13.0), Q10=(2.0, 1.0, 4.0)), [],
                                   hidden_layers=[12, 8], activation="sigmoid", scale_nn_outputs=True, input_batchnorm=True)


def m_two_chains_unequal_depth(eh, activation="tanh", bn=False):
    # hidden_layers as a NamedTuple per parameter (GenericHybridModel.jl:168-174): three hidden layers for rb, one for Q10
    return eh.constructHybridModel({"rb": ["sw_pot", "dsw_pot"], "Q10": ["ta"]}, ["ta"], ["reco"], eh.RbQ10,
                                   dict(rb=(3.0, 0.0, 13.0), Q10=(2.0, 1.0, 4.0)), [],
                                   hidden_layers={"rb": [16, 12, 8], "Q10": [10]}, activation=activation,
                                   scale_nn_outputs=True, input_batchnorm=bn)


def m_chains_mixed_activation(eh, acts=("tanh", "relu"), hidden=None, bn=True):
    # an activation per parameter (activation as a NamedTuple, GenericHybridModel.jl:168-174), here also with unequal depth
    hidden = hidden or {"rb": [12, 10], "Q10": [7]}
    return eh.constructHybridModel({"rb": ["sw_pot", "dsw_pot"], "Q10": ["ta", "sw_pot"]}, ["ta", "dsw_pot"], ["reco"], custom_pm,
                                   dict(rb=(3.0, 0.0, 13.0), Q10=(2.0, 1.0, 4.0), alpha=(0.5, -2.0, 2.0)), ["alpha"],
                                   hidden_layers=hidden, activation={"rb": acts[0], "Q10": acts[1]}, scale_nn_outputs=True,
                                   input_batchnorm=bn)


def m_traced_unequal_depth(eh):
    # chains of depth 1 and 2 (the shallow one first) feeding a traced process model with a global parameter besides; relu
    return eh.constructHybridModel({"rb": ["sw_pot", "dsw_pot"], "Q10": ["ta", "sw_pot"]}, ["ta", "dsw_pot"], ["reco"], custom_pm,
                                   dict(rb=(3.0, 0.0, 13.0), Q10=(2.0, 1.0, 4.0), alpha=(0.5, -2.0, 2.0)), ["alpha"],
                                   hidden_layers={"rb": [9], "Q10": [8, 7]}, activation="relu", scale_nn_outputs=True,
                                   input_batchnorm=True)   # (normalised inputs: relu into a saturating sigmoid is ill-conditioned in fp32)


CASES = [
    ("custom-tanh16", m_custom, lambda: _table(3000), "mse", "sum"),
    ("custom-nan", m_custom, lambda: _table(3000, nan_frac=0.05), "mae", "sum"),
    ("custom-relu32-noscale", lambda eh: m_custom(eh, hidden=(32, 32), activation="relu", scale=False), lambda: _table(2000), "mse", "sum"),
    ("custom-swish-bn-nse", lambda eh: m_custom(eh, hidden=(12, 12), activation="swish", bn=True), lambda: _table(2000), "nseLoss", "sum"),
    ("custom-two-neural", m_two_neural, lambda: _table(2000), "mse", "sum"),
    ("custom-two-targets-bn", m_two_targets, lambda: _table(2500, nan_frac=0.1, two=True), "PT", "mean"),
    ("rbq10-relu", lambda eh: rbq10_model(eh, activation="relu"), lambda: make_synth(2000), "rmse", "sum"),
    # one and three hidden layers (the specialised variants and the tensor-core path start at two)
    ("rbq10-one-hidden-layer", lambda eh: rbq10_model(eh, hidden=(16,)), lambda: make_synth(2000, nan_frac=0.03), "mse", "sum"),
    ("rbq10-three-hidden-layers", lambda eh: rbq10_model(eh, hidden=(16, 12, 8), activation="sigmoid"), lambda: make_synth(2000), "mse", "sum"),
    ("custom-three-hidden-two-neural", lambda eh: m_two_neural(eh, hidden=(32, 24, 20)), lambda: _table(2000), "mse", "sum"),
    ("custom-one-hidden-32-swish", lambda eh: m_custom(eh, hidden=(24,), activation="swish"), lambda: _table(2000), "mae", "sum"),
    ("rbq10-three-inputs-swish", m_rbq10_three_inputs, lambda: make_synth(2000, nan_frac=0.03), "mse", "sum"),
    # several chains, embedded block-diagonally into one chain of the summed widths (<= 32)
    ("two-chains-rbq10", lambda eh: rbq10_two_chain_model(eh, hidden=(16, 16)), lambda: make_synth(2000, nan_frac=0.03), "mse", "sum"),
    ("two-chains-six-inputs-bn", m_two_chains_six_inputs, lambda: make_synth(2000), "mse", "sum"),
    ("traced-all-operations", m_many_ops, lambda: _table(2000, nan_frac=0.03), "mse", "sum"),
    # chains of different depth: the shallower chain's last hidden layer rides on pass-through units
    ("two-chains-depth-3-and-1", m_two_chains_unequal_depth, lambda: make_synth(2000, nan_frac=0.03), "mse", "sum"),
    ("two-chains-depth-3-and-1-swish-bn", lambda eh: m_two_chains_unequal_depth(eh, "swish", True), lambda: make_synth(2000), "nseLoss", "sum"),
    ("traced-chains-depth-1-and-2-relu", m_traced_unequal_depth, lambda: _table(2000), "mse", "mean"),
]


def _setup(eh, orc, mk, mkdata, loss, agg, opt=None, seed=11):
    model = mk(eh)
    if loss == "PT":
        loss = eh.PerTarget("nseLoss", "mse")
    xf, y = eh.prepare_data(model, mkdata())
    rng = np.random.default_rng(seed)
    flat = model.initialparameters(rng)
    flat += (0.05 * rng.standard_normal(flat.size)).astype(np.float32)
    sess = eh.FusedSession(model, training_loss=loss, agg=agg, opt=opt)
    sess.upload(0, xf, y)
    sess.set_params(flat)
    o = orc.Oracle(model, training_loss=loss, agg=agg, opt=opt)
    return model, xf, y, flat, sess, o, rng


@pytest.mark.parametrize("name,mk,mkdata,loss,agg", CASES, ids=[c[0] for c in CASES])
def test_generic_variants_loss_and_gradient(eh, orc, name, mk, mkdata, loss, agg):
    model, xf, y, flat, sess, o, rng = _setup(eh, orc, mk, mkdata, loss, agg)
    assert sess.kernel_variant().startswith("ffma2/PmProgram/"), sess.kernel_variant()
    n = xf[0].shape[0]
    for B in (n, 517, 64, 12):
        idx = rng.permutation(n)[:B]
        L, g = sess.loss_grad(idx)
        L64, g64 = o.loss_grad(flat, xf, y, idx, precision=64)
        scale = np.abs(g64).max()
        assert abs(L - L64) <= RTOL_LOSS * abs(L64), (name, B, L, L64)
        err = np.abs(g - g64).max() / scale
        assert err <= RTOL_GRAD, (name, B, err)   # 1e-5 of the largest entry, no exception
    sess.close()


TRAIN_CASES = [c for c in CASES if c[0] in ("custom-tanh16", "custom-two-targets-bn", "rbq10-one-hidden-layer", "rbq10-three-hidden-layers",
                                            "custom-three-hidden-two-neural", "rbq10-three-inputs-swish", "two-chains-rbq10", "two-chains-six-inputs-bn",
                                            "traced-all-operations", "two-chains-depth-3-and-1", "two-chains-depth-3-and-1-swish-bn",
                                            "traced-chains-depth-1-and-2-relu")]


@pytest.mark.parametrize("name,mk,mkdata,loss,agg", TRAIN_CASES, ids=[c[0] for c in TRAIN_CASES])
def test_generic_variants_train_and_eval(eh, orc, name, mk, mkdata, loss, agg):
    """persistent epoch kernel, single steps and the eval kernel on the generic variants against the oracle"""
    model, xf, y, flat, sess, o, rng = _setup(eh, orc, mk, mkdata, loss, agg, opt=eh.Adam(0.01))
    n = xf[0].shape[0]
    B = 256
    perm = np.concatenate([rng.permutation(n) for _ in range(4)])[: 24 * B]
    got = sess.epoch(perm, B)
    ref = flat.copy()
    want = o.train_steps(ref, xf, y, perm, B)
    np.testing.assert_allclose(got, want, rtol=2e-4)
    ps = sess.get_params()
    ng = len(model.global_param_names)
    if ng:
        assert np.allclose(ps[-ng:], ref[-ng:], atol=2e-4), (ps[-ng:], ref[-ng:])   # phi (raw) after 24 Adam steps
    # single steps continue the same trajectory
    L1 = sess.step(perm[:B])
    Lo = o.train_steps(ref, xf, y, perm[:B], B)
    assert abs(L1 - Lo[0]) <= 2e-4 * abs(Lo[0])
    ps = sess.get_params()
    yhat, stats, par = sess.eval(0, want_yhat=True, want_params=True)
    want_y = o.forward(ps, xf, precision=64)
    assert yhat.shape == want_y.shape
    assert np.allclose(yhat, want_y, rtol=2e-5, atol=2e-5)
    _, want_p = o.forward(ps, xf, precision=64, want_params=True)
    rows = [list(model.parameters.names).index(nm) for nm in model.neural_param_names]
    assert np.allclose(par[rows], want_p[rows], rtol=2e-5, atol=2e-5)   # the neural parameters per sample
    sess.close()


def test_traced_model_trains_through_train_api(eh):
    """train() with a user-defined process model takes the fused fp32 path and recovers Q10 = 2 and alpha = 0.3"""
    table = _table(8192)
    model = m_custom(eh)
    res = eh.train(model, table, nepochs=30, batchsize=512, opt=eh.Adam(0.01), training_loss="mse", random_seed=1,
                   patience=100)
    assert res is not None
    last = res.val_history[-1]["mse"]["sum"]
    assert last < 0.03, last
    assert abs(res.train_diffs["Q10"] - 2.0) < 0.1, res.train_diffs["Q10"]
    assert abs(res.train_diffs["alpha"] - 0.3) < 0.15, res.train_diffs["alpha"]


def test_library_recognises_builtin_forms_in_traced_programs(eh, monkeypatch):
    """a host that only traces (the Julia shim) ships EH_PM_PROGRAM; eh_create matches it against the built-in forms and
    takes the specialised kernels -- same variant, same numbers as with the host-side matcher"""
    from conftest import expo_model, linear_model, make_synth, rbq10_model
    for mk, want in ((rbq10_model, "ffma2/PmRbQ10/"), (expo_model, "ffma2/PmExpo/"), (linear_model, "ffma2/PmLinear/")):
        monkeypatch.delenv("EH_PY_NO_MATCH", raising=False)
        a = eh.FusedSession(mk(eh))
        monkeypatch.setenv("EH_PY_NO_MATCH", "1")
        b = eh.FusedSession(mk(eh))
        assert a.kernel_variant() == b.kernel_variant() and b.kernel_variant().startswith(want), (a.kernel_variant(), b.kernel_variant())
        a.close(); b.close()
    # and a program that is NOT a built-in form stays a program
    def other(*, ta, Q10, rb):
        return {"reco": rb * np.exp(Q10) + ta}
    m = eh.constructHybridModel(["sw_pot"], ["ta"], ["reco"], other, dict(Q10=(2, 1, 4), rb=(3, 0, 13)), ["rb"], ["Q10"])
    s = eh.FusedSession(m)
    assert s.kernel_variant().startswith("ffma2/PmProgram/")
    s.close()
    # numbers: loss / gradient identical either way
    monkeypatch.delenv("EH_PY_NO_MATCH", raising=False)
    model = rbq10_model(eh)
    xf, y = eh.prepare_data(model, make_synth(1500))
    flat = model.initialparameters(np.random.default_rng(1))
    idx = np.arange(1000)
    res = []
    for env in (None, "1"):
        if env:
            monkeypatch.setenv("EH_PY_NO_MATCH", env)
        sess = eh.FusedSession(model)
        sess.upload(0, xf, y)
        sess.set_params(flat)
        res.append(sess.loss_grad(idx))
        sess.close()
    assert res[0][0] == res[1][0] and np.array_equal(res[0][1], res[1][1])
