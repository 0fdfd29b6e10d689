"""Host-side mirror of the reference API: constructors, descriptor, tracing, data prep and
splitting, config validation (the counterparts of test/test_generic_hybrid_model.jl:130-330,
test/test_split_data_train.jl:69-123 and test/test_loss_fn.jl:150-211 on the host logic)."""
import numpy as np
import pytest

from conftest import expo_model, linear_model, make_synth, rbq10_model


def test_constructor_dispatch_and_partition(eh):
    m = rbq10_model(eh)
    assert isinstance(m, eh.SingleNNHybridModel)
    assert m.neural_param_names == ["rb"] and m.global_param_names == ["Q10"] and m.fixed_param_names == []
    assert m.num_params() == 2 * 16 + 16 + 16 * 16 + 16 + 16 + 1 + 1
    m2 = expo_model(eh)
    assert isinstance(m2, eh.MultiNNHybridModel) and m2.neural_param_names == ["Resp0"]
    m3 = eh.constructHybridModel(["x2", "x3"], ["x1"], ["obs"], eh.LinearModel,
                                 dict(a=(1, 0, 5), b=(2, 0, 10), c=(0.5, 0, 2), d=(0.5, 0, 2)), ["a"], ["b"])
    assert m3.fixed_param_names == ["c", "d"]  # test_generic_hybrid_model.jl:153
    with pytest.raises(AssertionError):
        eh.constructHybridModel(["x2"], ["x1"], ["obs"], eh.LinearModel, dict(a=(1, 0, 5), b=(2, 0, 10)), ["zzz"], ["b"])
    with pytest.raises(TypeError):
        eh.constructHybridModel(3, ["x1"], ["obs"], eh.LinearModel, dict(a=(1, 0, 5), b=(2, 0, 10)), ["a"], ["b"])


def test_initial_parameters_layout(eh):
    m = rbq10_model(eh)
    flat = m.initialparameters(np.random.default_rng(0))
    assert flat.dtype == np.float32 and flat.size == m.num_params()
    tree = m.unflatten(flat)
    assert tree["ps"]["layer_1"]["weight"].shape == (16, 2) and tree["ps"]["layer_3"]["bias"].shape == (1,)
    assert np.all(tree["ps"]["layer_2"]["bias"] == 0)
    # phi starts at inv_sigmoid((default-l)/(u-l)) = logit(1/3)
    assert tree["Q10"][0] == pytest.approx(np.log((1 / 3) / (2 / 3)), abs=1e-6)
    # column-major weight order
    w = tree["ps"]["layer_1"]["weight"]
    assert flat[1] == w[1, 0] and flat[16] == w[0, 1]
    assert m.initialstates() == {"fixed": {}}


def test_process_model_tracing_and_builtin_match(eh):
    from easyhybrid_b200 import _abi
    d = eh.build_desc(rbq10_model(eh)).desc
    assert d.process_model == _abi.PM["RBQ10"] and d.pm_consts[0] == 15.0
    assert [(d.pm_args[i].kind, d.pm_args[i].index) for i in range(3)] == [(0, 0), (0, 1), (1, 0)]

    def rbq10_swapped(*, ta, Q10, rb, tref=10.0):  # same form, other constant / operand order
        return {"reco": (Q10 ** ((ta - tref) * 0.1)) * rb}
    m = eh.constructHybridModel(["sw_pot", "dsw_pot"], ["ta"], ["reco"], rbq10_swapped,
                                dict(Q10=(2, 1, 4), rb=(3, 0, 13)), ["rb"], ["Q10"])
    d = eh.build_desc(m).desc
    assert d.process_model == _abi.PM["RBQ10"] and d.pm_consts[0] == 10.0
    assert [(d.pm_args[i].kind, d.pm_args[i].index) for i in range(3)] == [(0, 1), (0, 0), (1, 0)]
    assert eh.build_desc(expo_model(eh)).desc.process_model == _abi.PM["EXPO"]
    assert eh.build_desc(linear_model(eh)).desc.process_model == _abi.PM["LINEAR"]
    assert eh.build_desc(linear_model(eh, two=True)).desc.process_model == _abi.PM["LINEAR2"]
    from conftest import expo2_model
    d2 = eh.build_desc(expo2_model(eh, hidden=(512, 512, 512))).desc
    assert d2.process_model == _abi.PM["EXPO2"] and d2.n_targ == 2 and d2.chains[0].n_hidden == 3

    def other(*, ta, Q10, rb):
        return {"reco": rb * np.exp(Q10) + np.sqrt(ta * ta)}
    m = eh.constructHybridModel(["sw_pot"], ["ta"], ["reco"], other, dict(Q10=(2, 1, 4), rb=(3, 0, 13)), ["rb"], ["Q10"])
    b = eh.build_desc(m)
    assert b.desc.process_model == _abi.PM["PROGRAM"] and b.desc.pm_len >= 6

    def bad(*, ta, Q10, rb, missing):
        return {"reco": rb}
    m = eh.constructHybridModel(["sw_pot"], ["ta"], ["reco"], bad, dict(Q10=(2, 1, 4), rb=(3, 0, 13)), ["rb"], ["Q10"])
    with pytest.raises(ValueError):
        eh.build_desc(m)


def test_prepare_data_drops_rows_like_reference(eh):
    m = rbq10_model(eh)
    t = make_synth(50)
    t["ta"][3] = np.nan          # NaN forcing -> dropped
    t["sw_pot"][7] = np.nan      # NaN predictor -> dropped
    t["reco"][11] = np.nan       # only target, all missing -> dropped
    (X, forc), targ = eh.prepare_data(m, t)
    assert X.shape == (47, 2) and X.dtype == np.float32 and forc["ta"].shape == (47,) and targ["reco"].shape == (47,)
    keep = np.setdiff1d(np.arange(50), [3, 7, 11])
    np.testing.assert_array_equal(X[:, 0], t["sw_pot"][keep])


def test_split_modes(eh):
    m = rbq10_model(eh)
    t = make_synth(100)
    tr, va, (itr, iva) = eh.split_data(t, m)
    assert tr[0][0].shape[0] == 80 and va[0][0].shape[0] == 20
    np.testing.assert_array_equal(itr, np.arange(80))          # splitobs(at=0.8, shuffle=false)
    np.testing.assert_array_equal(iva, np.arange(80, 100))
    tr, va, (itr, iva) = eh.split_data(t, m, shuffleobs=True, rng=np.random.default_rng(0))
    assert sorted(np.concatenate([itr, iva]).tolist()) == list(range(100)) and not np.array_equal(itr, np.arange(80))
    ids = np.repeat(np.arange(10), 10)
    tr, va, (itr, iva) = eh.split_data(t, m, split_by_id=ids)
    assert set(ids[itr]) == set(range(8)) and set(ids[iva]) == {8, 9}
    folds = np.tile(np.arange(1, 6), 20)
    tr, va, (itr, iva) = eh.split_data(t, m, folds=folds, val_fold=2)
    assert np.all(folds[iva] == 2) and len(itr) == 80
    with pytest.raises(ValueError):
        eh.split_data(t, m, split_by_id=ids, folds=folds, val_fold=1)
    with pytest.raises(AssertionError):
        eh.split_data(t, m, folds=folds)


def test_valid_mask_and_batches(eh):
    from easyhybrid_b200.data import batch_ranges, shard_batch
    masks, empty = eh.valid_mask({"a": np.array([1.0, np.nan]), "b": np.array([np.nan, np.nan])})
    assert masks["a"].tolist() == [True, False] and not empty
    assert eh.valid_mask({"a": np.array([np.nan])})[1]
    assert batch_ranges(10, 4) == [(0, 4), (4, 8), (8, 10)]
    parts = [shard_batch(100, 165, r, 4) for r in range(4)]
    assert parts[0][0] == 100 and parts[-1][1] == 165 and all(parts[i][1] == parts[i + 1][0] for i in range(3))


def test_config_validation(eh):
    eh.validate_config(eh.TrainConfig())
    for bad in (dict(return_model="x"), dict(batchsize=0), dict(nepochs=0), dict(patience=0)):
        with pytest.raises(ValueError):
            eh.validate_config(eh.TrainConfig(**bad))
    for metric in ("nse", "r2", "pearson", "kge"):   # check_training_loss, loss_fn.jl:196-205
        with pytest.raises(ValueError):
            eh.validate_config(eh.TrainConfig(training_loss=metric))
    assert eh.bestdirection("r2") == "Maximize" and eh.bestdirection("mse") == "Minimize"
    assert eh.isbetter(0.9, 0.8, "nse") and eh.isbetter(0.1, 0.2, "rmse")
    with pytest.raises(AssertionError):
        eh.build_desc(linear_model(eh, two=True), training_loss=eh.PerTarget("mse"))
