"""Pin the checker (and the GPU path) to the REAL reference: replay a dump written by julia/parity_dump.jl.

Julia is not part of this image, so the dump cannot be produced here ("parity unpinned" for Dense /
Zygote / Optimisers numerics).  Anyone with Julia and EasyHybrid.jl v0.2.0 closes the loop with

    julia --project=/path/to/EasyHybrid.jl julia/parity_dump.jl tests/golden/julia_dump.bin
    python -m pytest tests/test_julia_dump.py            # add -m gpu for the CUDA path

(or EH_JULIA_DUMP=/path/to/file).  Both tests skip while the file is absent.

Layout (little endian, julia/parity_dump.jl): Int64 n, B, nsteps, nflat; Float32 X[2, n] (column-major: one sample's two
predictors are contiguous), ta[n], reco[n]; Float32 ps0[nflat] (ComponentArray order); Int64 perm[nsteps * B] (1-based);
then per step: Float32 loss, Float32 grad[nflat], Float32 ps[nflat] (after the Adam(0.01) step).
The model is the README's RbQ10 hybrid (hidden [16, 16], tanh, scale_nn_outputs), training_loss :mse, agg sum:
Lux.Training.single_train_step!(AutoZygote(), ...) as called by src/training/epoch.jl:20-26."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
DUMP = os.environ.get("EH_JULIA_DUMP", os.path.join(HERE, "golden", "julia_dump.bin"))
needs_dump = pytest.mark.skipif(not os.path.exists(DUMP), reason="no Julia dump (julia/parity_dump.jl needs Julia + EasyHybrid.jl)")


def read_dump(path):
    raw = open(path, "rb").read()
    n, B, nsteps, nflat = (int(v) for v in np.frombuffer(raw, dtype="<i8", count=4))
    off = 32
    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(raw, dtype=dtype, count=count, offset=off)
        off += a.nbytes
        return a
    X = take("<f4", 2 * n).reshape(n, 2)
    ta, reco, ps0 = take("<f4", n), take("<f4", n), take("<f4", nflat).copy()
    perm = take("<i8", nsteps * B) - 1
    steps = []
    for _ in range(nsteps):
        steps.append((float(take("<f4", 1)[0]), take("<f4", nflat).astype(np.float64), take("<f4", nflat).copy()))
    assert off == len(raw), "trailing bytes: layout mismatch with julia/parity_dump.jl"
    return dict(n=n, B=B, nsteps=nsteps, nflat=nflat, X=X, ta=ta, reco=reco, ps0=ps0, perm=perm, steps=steps)


def _check_trajectory(d, loss_grad, step):
    """loss_grad(flat, idx) -> (L, g); step(idx) -> parameters after one optimiser step from the current state"""
    flat = d["ps0"].copy()
    B = d["B"]
    for k, (L_j, g_j, ps_j) in enumerate(d["steps"]):
        idx = d["perm"][k * B:(k + 1) * B]
        L, g = loss_grad(flat, idx)
        assert abs(L - L_j) <= 1e-5 * abs(L_j), (k, L, L_j)
        # Zygote's Float32 gradient against ours: 1e-4 of the max-norm (Float32 accumulation over 512 samples)
        assert np.abs(np.asarray(g, dtype=np.float64) - g_j).max() <= 1e-4 * np.abs(g_j).max(), k
        flat = step(idx)
        # Adam turns noise-level gradient entries into +-eta steps: compare phi tightly and theta in the bulk
        assert abs(float(flat[-1]) - float(ps_j[-1])) <= 1e-4, (k, flat[-1], ps_j[-1])
        assert np.median(np.abs(flat - ps_j)) <= 1e-5, k
        flat = ps_j.copy()   # re-anchor on the reference's parameters: every step is compared from identical state


def _model(eh):
    from conftest import rbq10_model
    return rbq10_model(eh)


@needs_dump
def test_checker_against_julia_dump(eh, orc):
    d = read_dump(DUMP)
    model = _model(eh)
    xf, y = (d["X"], {"ta": d["ta"]}), {"reco": d["reco"]}
    o = orc.Oracle(model, opt=eh.Adam(0.01))
    state = {"flat": d["ps0"].copy()}

    def loss_grad(flat, idx):
        state["flat"] = flat.copy()
        return o.loss_grad(flat, xf, y, idx, precision=32)

    def step(idx):
        f = state["flat"].copy()
        o.train_steps(f, xf, y, idx, len(idx))
        return f
    # the optimiser state of the checker follows the reference's step by step because every step starts from ps_j
    _check_trajectory(d, loss_grad, step)


@needs_dump
@pytest.mark.gpu
def test_gpu_against_julia_dump(eh):
    d = read_dump(DUMP)
    model = _model(eh)
    xf, y = (d["X"], {"ta": d["ta"]}), {"reco": d["reco"]}
    sess = eh.FusedSession(model, opt=eh.Adam(0.01))
    sess.upload(0, xf, y)

    def loss_grad(flat, idx):
        sess.set_params(flat)
        return sess.loss_grad(idx)

    def step(idx):
        sess.step(idx)
        return sess.get_params()
    _check_trajectory(d, loss_grad, step)
    sess.close()


def test_dump_layout_roundtrip(tmp_path, eh, orc):
    """the reader and the comparison logic above, exercised on a dump written HERE in the documented layout (by the
    checker standing in for Julia): guards the layout contract of julia/parity_dump.jl while no real dump exists"""
    from conftest import make_synth
    model = _model(eh)
    xf, y = eh.prepare_data(model, make_synth(4096))
    n = xf[0].shape[0]
    B, nsteps = 512, 4
    rng = np.random.default_rng(3)
    ps0 = model.initialparameters(rng)
    perm = rng.permutation(n)[: nsteps * B]
    o = orc.Oracle(model, opt=eh.Adam(0.01))
    path = tmp_path / "dump.bin"
    with open(path, "wb") as f:
        np.array([n, B, nsteps, ps0.size], dtype="<i8").tofile(f)
        np.ascontiguousarray(xf[0], dtype="<f4").tofile(f)
        np.asarray(xf[1]["ta"], dtype="<f4").tofile(f)
        np.asarray(y["reco"], dtype="<f4").tofile(f)
        ps0.astype("<f4").tofile(f)
        (perm + 1).astype("<i8").tofile(f)
        flat = ps0.copy()
        for k in range(nsteps):
            idx = perm[k * B:(k + 1) * B]
            L, g = o.loss_grad(flat, xf, y, idx, precision=32)
            o.train_steps(flat, xf, y, idx, B)
            np.array([L], dtype="<f4").tofile(f)
            np.asarray(g, dtype="<f4").tofile(f)
            flat.astype("<f4").tofile(f)
    d = read_dump(str(path))
    assert (d["n"], d["B"], d["nsteps"], d["nflat"]) == (n, B, nsteps, ps0.size)
    o2 = orc.Oracle(model, opt=eh.Adam(0.01))
    state = {}

    def loss_grad(fl, idx):
        state["flat"] = fl.copy()
        return o2.loss_grad(fl, (d["X"], {"ta": d["ta"]}), {"reco": d["reco"]}, idx, precision=64)

    def step(idx):
        f = state["flat"].copy()
        o2.train_steps(f, (d["X"], {"ta": d["ta"]}), {"reco": d["reco"]}, idx, len(idx))
        return f
    _check_trajectory(d, loss_grad, step)
