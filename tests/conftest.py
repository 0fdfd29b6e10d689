import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def make_synth(n=512, seed=42, nan_frac=0.0):
    """Q10-shaped synthetic table, the recipe of make_synth_df (test/test_split_data_train.jl:15-31).
    numpy's generator, not Julia's MersenneTwister: same distributions, different stream."""
    rng = np.random.default_rng(seed)
    ta = 10 + 10 * rng.standard_normal(n)
    sw_pot = np.abs(50 + 20 * rng.standard_normal(n))
    dsw_pot = np.concatenate([[0.0], np.diff(sw_pot)])
    true_rb = 3.0 + 0.02 * (sw_pot - sw_pot.mean())
    reco = true_rb * 2.0 ** (0.1 * (ta - 15.0)) + 0.1 * rng.standard_normal(n)
    if nan_frac > 0:
        reco = np.where(rng.random(n) < nan_frac, np.nan, reco)
    return {k: v.astype(np.float32) for k, v in dict(ta=ta, sw_pot=sw_pot, dsw_pot=dsw_pot, reco=reco).items()}


def make_expo(n=500, seed=2314):
    """projects/ExpoHybrid/ExpoHybridEstim.jl:39-47."""
    rng = np.random.default_rng(seed)
    T = rng.random(n) * 40 - 10
    SM = rng.random(n) * 0.8 + 0.1
    resp = 1.1 * np.exp(-8.0 * (SM - 0.6) ** 2) * np.exp(0.07 * T)
    obs = resp + rng.standard_normal(n) * 0.05 * resp.mean()
    return {k: v.astype(np.float32) for k, v in dict(T=T, SM=SM, Resp_obs=obs).items()}


def make_linear(n=1000, seed=123, two=False):
    """src/data/synthetic_test_data.jl:4-16 (gen_linear_data)."""
    rng = np.random.default_rng(seed)
    x = rng.random((n, 3)).astype(np.float32)
    a = np.exp(-5.0 * (x[:, 1] - 0.7) ** 2) + x[:, 2] / 10.0
    obs = a * x[:, 0] + 2.0 + 0.1 * rng.random()
    d = dict(x1=x[:, 0], x2=x[:, 1], x3=x[:, 2], obs=obs)
    if two:
        d = dict(x1=x[:, 0], x2=x[:, 1], x3=x[:, 2], var1=obs, var2=2 * a * x[:, 0] + 2.0 + 0.05 * rng.standard_normal(n))
    return {k: np.asarray(v, dtype=np.float32) for k, v in d.items()}


def rbq10_model(eh, hidden=(16, 16), activation="tanh", scale=True, bn=False):
    return eh.constructHybridModel(["sw_pot", "dsw_pot"], ["ta"], ["reco"], eh.RbQ10,
                                   dict(rb=(3.0, 0.0, 13.0), Q10=(2.0, 1.0, 4.0)), ["rb"], ["Q10"],
                                   hidden_layers=list(hidden), activation=activation, scale_nn_outputs=scale,
                                   input_batchnorm=bn)


def expo_model(eh, activation="sigmoid", scale=False, bn=False):
    return eh.constructHybridModel({"Resp0": ["SM"]}, ["T"], ["Resp_obs"], eh.Expo_resp_model,
                                   dict(k=(0.01, 0.0, 0.2), Resp0=(2.0, 0.0, 8.0)), ["k"],
                                   hidden_layers=[16, 16], activation=activation, scale_nn_outputs=scale,
                                   input_batchnorm=bn)


def make_expo2(n=500, seed=2314, nan_frac=0.0):
    """Expo recipe with a second target = 2 x first + noise (configuration C5)."""
    rng = np.random.default_rng(seed)
    T = rng.random(n) * 40 - 10
    SM = rng.random(n) * 0.8 + 0.1
    resp = 1.1 * np.exp(-8.0 * (SM - 0.6) ** 2) * np.exp(0.07 * T)
    obs = resp + rng.standard_normal(n) * 0.05 * resp.mean()
    obs2 = 2.0 * resp + rng.standard_normal(n) * 0.05 * resp.mean()
    if nan_frac:
        obs2[rng.random(n) < nan_frac] = np.nan
    return {k: v.astype(np.float32) for k, v in dict(T=T, SM=SM, Resp_obs=obs, Resp_obs2=obs2).items()}


def expo2_model(eh, hidden=(16, 16), activation="tanh", scale=False):
    return eh.constructHybridModel({"Resp0": ["SM"]}, ["T"], ["Resp_obs", "Resp_obs2"], eh.Expo_resp_model2,
                                   dict(k=(0.01, 0.0, 0.2), Resp0=(2.0, 0.0, 8.0)), ["k"],
                                   hidden_layers=list(hidden), activation=activation, scale_nn_outputs=scale)


def rbq10_two_chain_model(eh, hidden=(16, 16), activation="tanh"):
    """MultiNNHybridModel with two chains (src/models/GenericHybridModel.jl:169-189): rb and Q10 each get their own
    network over their own predictors"""
    return eh.constructHybridModel({"rb": ["sw_pot", "dsw_pot"], "Q10": ["ta"]}, ["ta"], ["reco"], eh.RbQ10,
                                   dict(rb=(3.0, 0.0, 13.0), Q10=(2.0, 1.0, 4.0)), [],
                                   hidden_layers=list(hidden), activation=activation, scale_nn_outputs=True)


def linear_model(eh, two=False, activation="relu"):
    fn = eh.LinearModel2 if two else eh.LinearModel
    targets = ["var1", "var2"] if two else ["obs"]
    return eh.constructHybridModel(["x2", "x3"], ["x1"], targets, fn, dict(a=(1.0, 0.0, 5.0), b=(2.0, 0.0, 10.0)),
                                   ["a"], ["b"], hidden_layers=[15, 15], activation=activation)


def truth_trajectory(o, flat0, xf, y, perm, B, nthreads=1):
    """per-step float64 loss of the oracle along its own trajectory: float64 loss / gradient, the oracle's Float32
    optimiser rule (Optimisers.jl semantics) -- a reference trajectory without float32 forward / backward error.
    All-masked batches are skipped (src/training/epoch.jl:17-19) and reported as NaN.  Returns (losses, final flat)."""
    ref = flat0.copy()
    want = []
    for k in range((perm.size + B - 1) // B):
        idx = perm[k * B:(k + 1) * B]
        if np.isnan(np.stack([y[t][idx] for t in o.model.targets])).all():
            want.append(np.nan)
            continue
        L, g = o.loss_grad(ref, xf, y, idx, precision=64, nthreads=nthreads)
        want.append(L)
        o.opt_step(ref, g.astype(np.float32))
    return np.array(want), ref


@pytest.fixture(scope="session")
def eh():
    import easyhybrid_b200
    return easyhybrid_b200


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.build()
    return oracle
