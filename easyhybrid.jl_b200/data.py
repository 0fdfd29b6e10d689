"""Data preparation and splitting: mirror of src/data/prepare_data.jl:3-63,
src/data/split_data.jl:8-79, 176-184, src/data/splits.jl:3-20 and valid_mask
(src/training/train.jl:221-232).  Index work only -- no model numerics."""
from __future__ import annotations

import numpy as np

from .model import predictor_columns


def _columns(data):
    if hasattr(data, "columns") and hasattr(data, "__getitem__") and not isinstance(data, dict):  # pandas
        return {str(c): np.asarray(data[c]) for c in data.columns}
    return {str(k): np.asarray(v) for k, v in dict(data).items()}


def prepare_data(hm, data, *, drop_missing_rows=True):
    """-> ((X [N, P] C-order == P x N column-major, forcings {name: [N]}), targets {name: [N]}),
    Float32, rows with a NaN predictor/forcing or no target at all dropped (prepare_data.jl:31-63)."""
    if isinstance(data, tuple):
        return data
    cols = _columns(data)
    pcols = predictor_columns(hm)
    need = list(dict.fromkeys(pcols + list(hm.forcing) + list(hm.targets)))
    for c in need:
        if c not in cols:
            raise KeyError(f"column `{c}` not found in data")
    f32 = {}
    for c in need:
        v = cols[c]
        if v.dtype == object:
            v = np.array([np.nan if x is None else x for x in v], dtype=np.float64)
        f32[c] = v.astype(np.float32)
    n = len(f32[need[0]])
    keep = np.ones(n, dtype=bool)
    if drop_missing_rows:
        for c in pcols + list(hm.forcing):
            keep &= ~np.isnan(f32[c])
        any_target = np.zeros(n, dtype=bool)
        for t in hm.targets:
            any_target |= ~np.isnan(f32[t])
        keep &= any_target
    X = np.ascontiguousarray(np.stack([f32[c][keep] for c in pcols], axis=1)) if pcols else np.zeros((int(keep.sum()), 0), np.float32)
    forc = {f: np.ascontiguousarray(f32[f][keep]) for f in hm.forcing}
    targ = {t: np.ascontiguousarray(f32[t][keep]) for t in hm.targets}
    return (X, forc), targ


def splitobs_indices(n, at=0.8, shuffle=False, rng=None):
    """MLUtils.splitobs(1:n; at, shuffle): first round(at*n) observations train, rest val
    (MLUtils 0.4.8, unpinned).  0-based indices."""
    idx = np.arange(n)
    if shuffle:
        idx = (rng or np.random.default_rng()).permutation(n)
    n1 = int(min(max(round(at * n), 0), n))
    return idx[:n1], idx[n1:]


def _take(xf_y, idx):
    (X, forc), targ = xf_y
    return (X[idx], {k: v[idx] for k, v in forc.items()}), {k: v[idx] for k, v in targ.items()}


def split_data(data, hm, *, split_by_id=None, folds=None, val_fold=None, shuffleobs=False, split_data_at=0.8,
               rng=None, **_ignored):
    """Three modes of split_data.jl:37-78: by id, external folds, plain splitobs."""
    cols = None if isinstance(data, tuple) else _columns(data)
    # ids / folds refer to rows of the raw table: carry them through the row filter
    prepared = prepare_data(hm, data)
    (X, forc), targ = prepared
    n = X.shape[0]
    if split_by_id is not None and folds is not None:
        raise ValueError("split_by_id and folds are not supported together; do the split when constructing folds")

    def aligned(v):
        v = cols[v] if isinstance(v, str) else np.asarray(v)
        if len(v) != n:
            raise AssertionError(f"length {len(v)} must equal number of samples ({n}); pass prepared data")
        return v

    if split_by_id is not None:
        ids = aligned(split_by_id)
        uniq = np.array(list(dict.fromkeys(ids.tolist())))
        tr_u, va_u = splitobs_indices(len(uniq), at=split_data_at, shuffle=shuffleobs, rng=rng)
        tr_ids, va_ids = set(uniq[tr_u].tolist()), set(uniq[va_u].tolist())
        train_idx = np.array([i for i, v in enumerate(ids.tolist()) if v in tr_ids], dtype=np.int64)
        val_idx = np.array([i for i, v in enumerate(ids.tolist()) if v in va_ids], dtype=np.int64)
    elif folds is not None or val_fold is not None:
        if val_fold is None:
            raise AssertionError("Provide val_fold when using folds.")
        if folds is None:
            raise AssertionError("Provide folds when using val_fold.")
        f = aligned(folds)
        if not (1 <= val_fold <= f.max()):
            raise AssertionError(f"val_fold={val_fold} is out of range 1:{f.max()}.")
        val_idx = np.nonzero(f == val_fold)[0]
        if len(val_idx) == 0:
            raise AssertionError(f"No samples assigned to validation fold {val_fold}.")
        train_idx = np.setdiff1d(np.arange(n), val_idx)
    else:
        train_idx, val_idx = splitobs_indices(n, at=split_data_at, shuffle=shuffleobs, rng=rng)
    return _take(prepared, train_idx), _take(prepared, val_idx), (train_idx, val_idx)


def valid_mask(y):
    """train.jl:221-232: per-target !isnan mask and whether everything is masked."""
    masks = {k: ~np.isnan(v) for k, v in y.items()}
    return masks, not any(m.any() for m in masks.values())


def batch_ranges(n, batchsize):
    """DataLoader batches: consecutive slices of the permutation, last one partial."""
    return [(a, min(a + batchsize, n)) for a in range(0, n, batchsize)]


def shard_batch(a, b, rank, world):
    """Data-parallel slice of a global batch [a, b): rank r owns the contiguous part
    [a + r*len/W, a + (r+1)*len/W)."""
    ln = b - a
    return a + (rank * ln) // world, a + ((rank + 1) * ln) // world
