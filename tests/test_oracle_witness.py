"""Oracle gradients / losses / optimiser rules cross-witnessed by an independent torch-CPU float64
implementation (tests/witness.py).  The reference's own tests pin none of these numerically
(SURVEY.md section 4), so this is the strongest check available without Julia."""
import numpy as np
import pytest
import torch

from conftest import (expo2_model, expo_model, linear_model, make_expo, make_expo2, make_linear, make_synth, rbq10_model,
                      rbq10_two_chain_model)
from witness import objective


def _prep(eh, model, table):
    (xf, y) = eh.prepare_data(model, table)
    return xf, y


CASES = [
    ("rbq10-tanh", lambda eh: rbq10_model(eh), lambda: make_synth(300), "mse", "sum"),
    ("rbq10-swish-nan", lambda eh: rbq10_model(eh, activation="swish"), lambda: make_synth(300, nan_frac=0.1), "mse", "sum"),
    ("rbq10-sigmoid-noscale", lambda eh: rbq10_model(eh, activation="sigmoid", scale=False), lambda: make_synth(200), "rmse", "sum"),
    ("rbq10-32-mae", lambda eh: rbq10_model(eh, hidden=(32, 32)), lambda: make_synth(200), "mae", "sum"),
    ("rbq10-bn", lambda eh: rbq10_model(eh, bn=True), lambda: make_synth(256), "mse", "sum"),
    ("expo-nse", lambda eh: expo_model(eh), lambda: make_expo(300), "nseLoss", "sum"),
    ("expo-bn-nse", lambda eh: expo_model(eh, bn=True), lambda: make_expo(300), "nseLoss", "sum"),
    ("linear-relu", lambda eh: linear_model(eh), lambda: make_linear(400), "mse", "sum"),
    ("linear2-mean", lambda eh: linear_model(eh, two=True, activation="tanh"), lambda: make_linear(400, two=True), "mse", "mean"),
    ("linear2-pertarget", lambda eh: linear_model(eh, two=True, activation="tanh"), lambda: make_linear(400, two=True), "PT", "sum"),
    # the two-target Expo form of the wide configuration (C5), here with a CPU-sized chain of three hidden layers
    # two Dense chains (MultiNNHybridModel with a network per parameter)
    ("rbq10-two-chains", lambda eh: rbq10_two_chain_model(eh), lambda: make_synth(300, nan_frac=0.05), "mse", "sum"),
    ("expo2-3x48-pertarget-nan", lambda eh: expo2_model(eh, hidden=(48, 48, 48)), lambda: make_expo2(300, nan_frac=0.1), "PT", "sum"),
]

# the models the generic exact-fp32 GPU variants are checked with (tests/test_gpu_generic.py): traced process models,
# one / three hidden layers, several chains -- the oracle is their checker on the GPU, this is the oracle's own witness
import test_gpu_generic as gg  # noqa: E402

CASES += [
    ("traced-custom", gg.m_custom, lambda: gg._table(300), "mse", "sum"),
    ("traced-custom-nan-mae", gg.m_custom, lambda: gg._table(300, nan_frac=0.05), "mae", "sum"),
    ("traced-relu32-noscale", lambda eh: gg.m_custom(eh, hidden=(32, 32), activation="relu", scale=False), lambda: gg._table(300), "mse", "sum"),
    ("traced-swish-bn-nse", lambda eh: gg.m_custom(eh, hidden=(12, 12), activation="swish", bn=True), lambda: gg._table(300), "nseLoss", "sum"),
    ("traced-two-neural", gg.m_two_neural, lambda: gg._table(300), "mse", "sum"),
    ("traced-two-targets-bn", gg.m_two_targets, lambda: gg._table(400, nan_frac=0.1, two=True), "PT", "mean"),
    # losses whose seeds depend on statistics of the predictions (forward pre-pass): loss_fn.jl:75-77, 104-127, 160-174
    ("rbq10-pearson", lambda eh: rbq10_model(eh), lambda: make_synth(300, nan_frac=0.05), "pearsonLoss", "sum"),
    ("rbq10-kge", lambda eh: rbq10_model(eh), lambda: make_synth(300), "kgeLoss", "sum"),
    ("expo-pbkge-bn", lambda eh: expo_model(eh, bn=True), lambda: make_expo(300), "pbkgeLoss", "sum"),
    ("linear2-rmse-two-targets", lambda eh: linear_model(eh, two=True, activation="tanh"), lambda: make_linear(400, two=True), "rmse", "mean"),
    ("linear2-kge-mse", lambda eh: linear_model(eh, two=True, activation="tanh"), lambda: make_linear(400, two=True), "PT-kge-mse", "sum"),
    ("rbq10-one-hidden-layer", lambda eh: rbq10_model(eh, hidden=(16,)), lambda: make_synth(300, nan_frac=0.03), "mse", "sum"),
    ("rbq10-three-hidden-layers", lambda eh: rbq10_model(eh, hidden=(16, 12, 8), activation="sigmoid"), lambda: make_synth(300), "mse", "sum"),
    ("rbq10-three-inputs-swish", gg.m_rbq10_three_inputs, lambda: make_synth(300, nan_frac=0.03), "mse", "sum"),
    ("two-chains-six-inputs-bn", gg.m_two_chains_six_inputs, lambda: make_synth(300), "mse", "sum"),
    ("traced-all-operations", gg.m_many_ops, lambda: gg._table(300, nan_frac=0.03), "mse", "sum"),
    ("two-chains-depth-3-and-1", gg.m_two_chains_unequal_depth, lambda: make_synth(300, nan_frac=0.03), "mse", "sum"),
    ("two-chains-depth-3-and-1-swish-bn", lambda eh: gg.m_two_chains_unequal_depth(eh, "swish", True), lambda: make_synth(300), "nseLoss", "sum"),
    ("traced-chains-depth-1-and-2-relu", gg.m_traced_unequal_depth, lambda: gg._table(300), "mse", "mean"),
]
# shapes the GPU only reaches with run-time compilation (tests/test_jit.py): three chain outputs, ten inputs
import test_jit as tj  # noqa: E402

CASES += [
    ("three-neural-parameters-width-24", tj.m_three_neural, lambda: gg._table(300, nan_frac=0.03), "mse", "sum"),
    ("ten-inputs", tj.m_ten_inputs, lambda: tj._table_wide(300), "nseLoss", "sum"),
    ("chains-tanh-and-relu-depth-2-and-1", gg.m_chains_mixed_activation, lambda: gg._table(300, nan_frac=0.03), "mse", "sum"),
    ("chains-swish-and-sigmoid", lambda eh: gg.m_chains_mixed_activation(eh, ("swish", "sigmoid"), {"rb": [9, 9], "Q10": [6, 5]}),
     lambda: gg._table(300), "mae", "sum"),
]


@pytest.mark.parametrize("agg,branches,normalize", [("sum", None, True), ("mean", None, False), ("sum", ["Q10"], True)])
def test_weight_l2_extra_loss_vs_autograd(eh, orc, agg, branches, normalize):
    """extra_loss = (ŷ, ps) -> (; l2 = λ * weight_l2(ps.<branch>; normalize),) (extract_weights.jl:55-91), loss = agg([L, l2])
    (compute_loss.jl:31-34), in its native declarative form WeightL2"""
    from conftest import rbq10_two_chain_model
    model = rbq10_two_chain_model(eh) if branches else rbq10_model(eh)
    xl = eh.WeightL2(0.3, branches=branches, normalize=normalize)
    xf, y = _prep(eh, model, make_synth(300, nan_frac=0.05))
    rng = np.random.default_rng(3)
    flat = model.initialparameters(rng)
    flat += (0.1 * rng.standard_normal(flat.size)).astype(np.float32)
    idx = rng.permutation(xf[0].shape[0])[:200]
    o = orc.Oracle(model, training_loss="mse", agg=agg, extra_loss=xl)
    L, g = o.loss_grad(flat, xf, y, idx, precision=64)
    Lw, gw = objective(model, flat, xf, y, idx, training_loss="mse", agg=agg, extra_loss=xl)
    L0, _ = orc.Oracle(model, training_loss="mse", agg=agg).loss_grad(flat, xf, y, idx, precision=64)
    assert abs(L - Lw) <= 1e-10 * abs(Lw) and np.abs(g - gw).max() <= 1e-9 * np.abs(gw).max()
    # and the host-side value (evaluation-mode bookkeeping) is the same term
    w2 = 0.5 if agg == "mean" else 1.0
    assert abs(L - w2 * (L0 + xl.value(model, flat))) <= 1e-6 * abs(L)


@pytest.mark.parametrize("name,mk,mkdata,loss,agg", CASES, ids=[c[0] for c in CASES])
def test_loss_and_grad_vs_autograd(eh, orc, name, mk, mkdata, loss, agg):
    model = mk(eh)
    if loss == "PT":
        loss = eh.PerTarget("nseLoss", "mse")
    if loss == "PT-kge-mse":
        loss = eh.PerTarget("kgeLoss", "mse")
    xf, y = _prep(eh, model, mkdata())
    rng = np.random.default_rng(7)
    flat = model.initialparameters(rng)
    flat += (0.05 * rng.standard_normal(flat.size)).astype(np.float32)  # non-zero biases, phi off default
    n = xf[0].shape[0]
    idx = rng.permutation(n)[: n - 37]
    o = orc.Oracle(model, training_loss=loss, agg=agg)
    Lw, gw = objective(model, flat, xf, y, idx, training_loss=loss, agg=agg)
    L64, g64 = o.loss_grad(flat, xf, y, idx, precision=64)
    L32, g32 = o.loss_grad(flat, xf, y, idx, precision=32, nthreads=2)
    scale = np.abs(gw).max()
    assert L64 == pytest.approx(Lw, rel=1e-10)
    assert np.abs(g64 - gw).max() <= 1e-9 * scale
    assert L32 == pytest.approx(Lw, rel=2e-6)
    assert np.abs(g32 - gw).max() <= 1e-5 * scale  # the tolerance the north star states for fp32


def test_thread_count_does_not_change_results(eh, orc):
    model = rbq10_model(eh)
    xf, y = _prep(eh, model, make_synth(1000))
    flat = model.initialparameters(np.random.default_rng(1))
    o = orc.Oracle(model)
    a = o.loss_grad(flat, xf, y, np.arange(1000), nthreads=1)
    b = o.loss_grad(flat, xf, y, np.arange(1000), nthreads=4)
    assert a[0] == b[0] and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("optname", ["Adam", "AdamW", "RMSProp", "Descent"])
def test_optimiser_rules_vs_torch(eh, orc, optname):
    """Optimisers.jl rules against torch.optim on a fixed gradient sequence"""
    model = rbq10_model(eh)
    opt = {"Adam": eh.Adam(0.01), "AdamW": eh.AdamW(0.01, (0.9, 0.999), 0.01), "RMSProp": eh.RMSProp(0.001),
           "Descent": eh.Descent(0.05)}[optname]
    o = orc.Oracle(model, opt=opt)
    rng = np.random.default_rng(5)
    flat = rng.standard_normal(o.n_flat).astype(np.float32)
    th = torch.tensor(flat.astype(np.float64), requires_grad=True)
    topt = {"Adam": lambda: torch.optim.Adam([th], lr=0.01, eps=1e-8),
            "AdamW": lambda: torch.optim.AdamW([th], lr=0.01, eps=1e-8, weight_decay=0.01),
            "RMSProp": lambda: torch.optim.RMSprop([th], lr=0.001, alpha=0.9, eps=1e-8),
            "Descent": lambda: torch.optim.SGD([th], lr=0.05)}[optname]()
    for _ in range(25):
        g = rng.standard_normal(o.n_flat).astype(np.float32)
        o.opt_step(flat, g)
        th.grad = torch.tensor(g.astype(np.float64))
        topt.step()
    # torch's AdamW decays before the Adam step (theta*(1-lr*wd)); Optimisers adds eta*lambda*theta to the
    # update: identical to first order in lr*wd
    tol = 2e-5 if optname == "AdamW" else 2e-6
    np.testing.assert_allclose(flat, th.detach().numpy(), rtol=0, atol=tol * max(1.0, np.abs(flat).max()))


def test_train_steps_skip_all_masked_batch(eh, orc):
    """run_epoch! skips a batch whose targets are all NaN (src/training/epoch.jl:17-19)"""
    model = rbq10_model(eh)
    table = make_synth(128)
    table["reco"][:32] = np.nan
    xf, y = (np.stack([table["sw_pot"], table["dsw_pot"]], 1), {"ta": table["ta"]}), {"reco": table["reco"]}
    o = orc.Oracle(model)
    flat = model.initialparameters(np.random.default_rng(2))
    before = flat.copy()
    losses = o.train_steps(flat, xf, y, np.arange(128), 32)
    assert np.isnan(losses[0]) and not np.isnan(losses[1:]).any()
    assert o.t.value == 3 and not np.array_equal(before, flat)
