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


JIT_CASES = [
    ("custom-tanh16", gg.m_custom, lambda: gg._table(3000, nan_frac=0.05), "mse", "sum"),
    ("custom-two-targets-bn", gg.m_two_targets, lambda: gg._table(2500, nan_frac=0.1, two=True), "PT", "mean"),
    ("traced-chains-depth-1-and-2-relu", gg.m_traced_unequal_depth, lambda: gg._table(2000), "mse", "mean"),
]


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
    si = eh.FusedSession(model, training_loss=loss, agg=agg, opt=eh.Adam(0.01))
    assert sj.kernel_variant().startswith("nvrtc/PmTraced#"), sj.kernel_variant()
    assert si.kernel_variant().startswith("ffma2/PmProgram/"), si.kernel_variant()
    n = xf[0].shape[0]
    for s in (sj, si):
        s.upload(0, xf, y)
        s.set_params(flat)
    for B in (n, 517, 12):
        idx = rng.permutation(n)[:B]
        L, g = sj.loss_grad(idx)
        Li, gi = si.loss_grad(idx)
        L64, g64 = o.loss_grad(flat, xf, y, idx, precision=64)
        scale = np.abs(g64).max()
        assert abs(L - L64) <= 2e-6 * abs(L64), (name, B, L, L64)
        assert np.abs(g - g64).max() <= 1e-5 * scale, (name, B)
        assert np.abs(g - gi).max() <= 2e-6 * scale and abs(L - Li) <= 1e-6 * abs(Li)   # same formulas, contraction aside
    # persistent epoch kernel, single step, eval
    B = 256
    perm = np.concatenate([rng.permutation(n) for _ in range(3)])[: 20 * B]
    got = sj.epoch(perm, B)
    ref = flat.copy()
    want = o.train_steps(ref, xf, y, perm, B)
    np.testing.assert_allclose(got, want, rtol=2e-4)
    np.testing.assert_allclose(got, si.epoch(perm, B), rtol=2e-5)
    L1 = sj.step(perm[:B])
    Lo = o.train_steps(ref, xf, y, perm[:B], B)
    assert abs(L1 - Lo[0]) <= 2e-4 * abs(Lo[0])
    ps = sj.get_params()
    yhat, stats, par = sj.eval(0, want_yhat=True, want_params=True)
    want_y = o.forward(ps, xf, precision=64)
    assert np.allclose(yhat, want_y, rtol=2e-5, atol=2e-5)
    sj.close()
    si.close()


@pytest.mark.gpu
def test_train_api_with_compiled_program(eh):
    table = gg._table(8192)
    res = eh.train(gg.m_custom(eh), table, nepochs=30, batchsize=512, opt=eh.Adam(0.01), training_loss="mse", random_seed=1,
                   patience=100, jit=True)
    assert res.val_history[-1]["mse"]["sum"] < 0.03
    assert abs(res.train_diffs["Q10"] - 2.0) < 0.1
