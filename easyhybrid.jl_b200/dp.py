"""Data-parallel host logic: how a global batch is split over ranks and how per-rank
results combine.  The device side does the same sums inside the persistent kernel over NVLink peer
memory; these helpers are what the tests use to state the contract (and to check it on CPU with gloo)."""
from __future__ import annotations

import numpy as np


def global_batch_indices(perms, sizes, batchsize, k):
    """indices (into the concatenation of the rank shards) of global batch k = union of the ranks' local
    batches k; ``perms[r]`` is rank r's local permutation, ``sizes[r]`` its shard size"""
    offs = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    parts = [offs[r] + np.asarray(p[k * batchsize:(k + 1) * batchsize]) for r, p in enumerate(perms)]
    return np.concatenate(parts)


def combine_mse(n_valid, loss, grad):
    """global mse loss / gradient from per-rank values computed with their LOCAL 1/n scaling:
    L = sum n_r L_r / sum n_r, g = sum n_r g_r / sum n_r.  Inputs are arrays over ranks."""
    n = np.asarray(n_valid, dtype=np.float64)
    w = n / n.sum()
    return float((w * np.asarray(loss, dtype=np.float64)).sum()), (w[:, None] * np.asarray(grad, dtype=np.float64)).sum(0)
