"""train(): mirror of the Lux.Training route of src/training/train.jl:95-136, 211-219,
run_epoch! (src/training/epoch.jl:13-33), evaluate_epoch (:53-66), early stopping
(src/training/early_stopping.jl:16-71) and TrainResults (TrainingConfig.jl:190-223).
All model arithmetic runs in libeasyhybrid_cuda.so through FusedSession; this file is the
epoch loop, index streams and bookkeeping that the reference keeps on the host as well."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .config import DataConfig, TrainConfig, is_optimisers_rule, override_configs, validate_config
from .data import split_data, valid_mask
from .losses import assemble_losses, isbetter
from .session import FusedSession


@dataclass
class EpochSnapshot:
    """initialization.jl:53-58."""
    l_train: dict
    l_val: dict
    yhat_train: object = None
    yhat_val: object = None


@dataclass
class TrainResults:
    train_history: list = field(default_factory=list)
    val_history: list = field(default_factory=list)
    step_losses: list = field(default_factory=list)
    ps: np.ndarray = None
    ps_tree: dict = None
    st: dict = None
    train_obs_pred: dict = None
    val_obs_pred: dict = None
    train_diffs: dict = None
    val_diffs: dict = None
    best_epoch: int = 0
    best_loss: float = float("nan")
    opt_state: tuple = None
    split_indices: tuple = None


def evaluate_epoch(sess, model, cfg, want_pred=False):
    """evaluate_acc on train and val (test mode), all cfg.loss_types."""
    out = []
    for split in (0, 1):
        if sess.n[split] == 0:
            out.append(({}, None, None))
            continue
        yhat, stats, par = sess.eval(split, want_yhat=want_pred, want_params=want_pred)
        losses = assemble_losses(stats, model.targets, list(dict.fromkeys(list(cfg.loss_types) + [_tl_name(cfg)])), cfg.agg)
        out.append((losses, yhat, par))
    return out


def _tl_name(cfg):
    tl = cfg.training_loss
    return str(tl) if isinstance(tl, str) else str(tl.losses[0])


def _extract_agg_loss(l_val, name, cfg):
    """extract_agg_loss(l_val) (early_stopping.jl:47-49): the aggregate of the first loss type; the reference reads the
    `sum` field whatever cfg.agg is -- with another aggregation the mirror reads that aggregate instead of failing"""
    entry = l_val[name]
    return entry["sum"] if "sum" in entry else entry[_agg_name(cfg)]


def _agg_name(cfg):
    return cfg.agg if isinstance(cfg.agg, str) else getattr(cfg.agg, "__name__", "sum")


def run_epoch(sess, perm0, cfg):
    """run_epoch!: one optimiser step per DataLoader batch, all-masked batches skipped."""
    return sess.epoch(perm0, cfg.batchsize)


def train(model, data, save_ps=(), *, train_cfg=None, data_cfg=None, jit=False, **kwargs):
    """train(model, data; kwargs...) -- flat kwargs override TrainConfig / DataConfig fields
    (train.jl:211-219, 300-314).  Returns TrainResults, or None when a split is empty."""
    train_cfg, data_cfg, rest = override_configs(train_cfg or TrainConfig(), data_cfg or DataConfig(), kwargs)
    if rest:
        import warnings
        warnings.warn(f"Unknown kwargs ignored on the Optimisers.jl path: {', '.join(rest)}")
    if not is_optimisers_rule(train_cfg.opt):
        raise NotImplementedError("only Optimisers.jl rules (Adam, AdamW, RMSProp, Descent) take the fused CUDA path")
    from .model import WeightL2
    if train_cfg.extra_loss is not None and not isinstance(train_cfg.extra_loss, WeightL2):
        raise NotImplementedError("extra_loss closures cannot cross the C ABI (EH_EUNSUPPORTED); "
                                  "WeightL2(lam, branches, normalize) is the native form of lambda * weight_l2(ps.<branch>)")
    cfg = validate_config(train_cfg)
    rng = np.random.default_rng(cfg.random_seed)  # seed! before split, loader and init (train.jl:98)

    (xf_tr, y_tr), (xf_va, y_va), split_idx = split_data(
        data, model, split_by_id=data_cfg.split_by_id, folds=data_cfg.folds, val_fold=data_cfg.val_fold,
        shuffleobs=data_cfg.shuffleobs, split_data_at=data_cfg.split_data_at, rng=rng)
    n_tr, n_va = xf_tr[0].shape[0], xf_va[0].shape[0]
    if n_tr == 0 or n_va == 0:
        return None
    _, all_masked = valid_mask(y_tr)

    device = cfg.gdev if isinstance(cfg.gdev, int) else 0
    sess = FusedSession(model, training_loss=cfg.training_loss, agg=cfg.agg, opt=cfg.opt, device=device, extra_loss=cfg.extra_loss, jit=jit)
    try:
        sess.upload(0, xf_tr, y_tr)
        sess.upload(1, xf_va, y_va)
        if cfg.train_from is None:
            ps0 = model.initialparameters(rng)
        else:
            tf = cfg.train_from
            ps0 = np.asarray(tf.ps if isinstance(tf, TrainResults) else tf[0], dtype=np.float32)
        sess.set_params(ps0)

        res = TrainResults(split_indices=split_idx)
        (l_tr, _, _), (l_va, _, _) = evaluate_epoch(sess, model, cfg)
        res.train_history.append(l_tr)
        res.val_history.append(l_va)
        # early stopping tracks the FIRST entry of cfg.loss_types, aggregated over the targets (extract_agg_loss reads the
        # `sum` field of l_val[1], early_stopping.jl:47-49) -- not the training loss
        es_name = str(cfg.loss_types[0])
        bn_on = bool(model.chains and model.chains[0]["input_batchnorm"])
        get_st = (lambda: [sess.get_bn_state(k) for k in range(len(model.chains))]) if bn_on else (lambda: None)
        best_loss, best_epoch, best_ps, best_st, wait = _extract_agg_loss(l_va, es_name, cfg), 0, ps0.copy(), get_st(), 0
        for epoch in range(1, cfg.nepochs + 1):
            perm0 = rng.permutation(n_tr)
            if not all_masked:
                res.step_losses.append(run_epoch(sess, perm0, cfg))
            (l_tr, _, _), (l_va, _, _) = evaluate_epoch(sess, model, cfg)
            if cfg.keep_history:
                res.train_history.append(l_tr)
                res.val_history.append(l_va)
            else:
                res.train_history, res.val_history = [l_tr], [l_va]
            cur = _extract_agg_loss(l_va, es_name, cfg)
            if isbetter(cur, best_loss, es_name):
                best_loss, best_epoch, best_ps, best_st, wait = cur, epoch, sess.get_params(), get_st(), 0
            else:
                wait += 1
            if wait >= cfg.patience:
                break
        final_ps = sess.get_params()
        if cfg.return_model == "best":
            # best_or_final (early_stopping.jl:51-71): stopper.best_ps / best_st -- the INITIAL parameters and states
            # when no epoch improved on them (best_epoch == 0)
            sess.set_params(best_ps)
            if bn_on:
                for k, (mu, var) in enumerate(best_st):
                    sess.set_bn_state(mu, var, k)
            res.ps = best_ps
        else:
            res.ps = final_ps
        res.best_epoch, res.best_loss = best_epoch, best_loss
        res.opt_state = sess.get_opt_state()
        res.ps_tree = model.unflatten(res.ps)
        res.st = model.initialstates()
        if bn_on:
            # running statistics of every chain's input BatchNorm (Lux `st`), chain after chain
            bn = [sess.get_bn_state(k) for k in range(len(model.chains))]
            res.st["bn"] = bn[0] if len(bn) == 1 else bn
        (_, yh_tr, par_tr), (_, yh_va, par_va) = evaluate_epoch(sess, model, cfg, want_pred=True)
        names = model.parameters.names
        res.train_obs_pred = {t: (y_tr[t], yh_tr[i]) for i, t in enumerate(model.targets)}
        res.val_obs_pred = {t: (y_va[t], yh_va[i]) for i, t in enumerate(model.targets)}
        glob = {g: float(model.parameters.column(1)[g] + (model.parameters.column(2)[g] - model.parameters.column(1)[g])
                         / (1.0 + np.exp(-float(res.ps[len(res.ps) - len(model.global_param_names) + j]))))
                for j, g in enumerate(model.global_param_names)}
        res.train_diffs = {**{n: par_tr[names.index(n)] for n in model.neural_param_names}, **glob}
        res.val_diffs = {**{n: par_va[names.index(n)] for n in model.neural_param_names}, **glob}
        return res
    finally:
        sess.close()
