"""FusedSession: one ``eh_ctx`` of libeasyhybrid_cuda.so, wrapped for the Python host.

Every method is a thin ctypes call through the C ABI (include/easyhybrid_cuda.h); numpy
arrays are only marshalled.  There is no CPU code path behind these calls."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from ._lib import EasyHybridCudaError, load
from .model import build_desc, predictor_columns

_fp = C.POINTER(C.c_float)


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_fp)


def _ptr_array(arrs):
    keep = [np.ascontiguousarray(a, dtype=np.float32) for a in arrs]
    pa = (_fp * max(len(keep), 1))(*[a.ctypes.data_as(_fp) for a in keep])
    return keep, pa


class FusedSession:
    def __init__(self, model, *, training_loss="mse", agg="sum", opt=None, device=0, flags=0, extra_loss=None, jit=False):
        """jit=True: a traced process model is compiled into the kernels at creation (NVRTC, EH_FLAG_JIT; seconds the first
        time, cached on disk afterwards) instead of being interpreted per sample."""
        self.lib = load()
        self.model = model
        if jit:
            flags |= _abi.EH_FLAG_JIT
        self.bundle = build_desc(model, training_loss=training_loss, agg=agg, opt=opt, device=device, flags=flags, extra_loss=extra_loss)
        h = C.c_void_p()
        st = self.lib.eh_create(C.byref(h), self.bundle.byref())
        if st != _abi.EH_OK:
            raise EasyHybridCudaError(st, (self.lib.eh_last_error(None) or b"").decode())
        self.h = h
        self.n_flat = int(self.lib.eh_num_params(self.h))
        self.pcols = predictor_columns(model)
        self.n = {0: 0, 1: 0}

    def _ck(self, st):
        if st != _abi.EH_OK:
            raise EasyHybridCudaError(st, (self.lib.eh_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.eh_destroy(self.h)
            self.h = None
            for p in getattr(self, "_pinned", []):
                self.lib.eh_host_free(p)
            self._pinned = []

    __del__ = close

    # ---- data ----
    def upload(self, split, xf, y):
        X, forc = xf
        Xc, Xp = _f32(X)
        kf, pf = _ptr_array([forc[f] for f in self.model.forcing])
        kt, pt = _ptr_array([y[t] for t in self.model.targets])
        n = Xc.shape[0]
        self._ck(self.lib.eh_upload(self.h, split, n, Xp, pf, pt))
        self.n[split] = n

    # ---- parameters / state ----
    def set_params(self, flat):
        a, p = _f32(flat)
        self._ck(self.lib.eh_set_params(self.h, p, a.size))

    def get_params(self):
        out = np.empty(self.n_flat, dtype=np.float32)
        self._ck(self.lib.eh_get_params(self.h, out.ctypes.data_as(_fp), out.size))
        return out

    def set_opt_state(self, m=None, v=None, t=0):
        km, pm = _f32(m) if m is not None else (None, None)
        kv, pv = _f32(v) if v is not None else (None, None)
        self._ck(self.lib.eh_set_opt_state(self.h, pm, pv, self.n_flat, int(t)))

    def get_opt_state(self):
        m = np.empty(self.n_flat, dtype=np.float32)
        v = np.empty(self.n_flat, dtype=np.float32)
        t = C.c_int64(0)
        self._ck(self.lib.eh_get_opt_state(self.h, m.ctypes.data_as(_fp), v.ctypes.data_as(_fp), self.n_flat, C.byref(t)))
        return m, v, int(t.value)

    def get_bn_state(self, chain=0):
        n = len(self.model.chains[chain]["predictors"])
        mean, var = np.empty(n, np.float32), np.empty(n, np.float32)
        self._ck(self.lib.eh_get_bn_state(self.h, chain, mean.ctypes.data_as(_fp), var.ctypes.data_as(_fp), n))
        return mean, var

    def set_bn_state(self, mean, var, chain=0):
        a, pa = _f32(mean)
        b, pb = _f32(var)
        self._ck(self.lib.eh_set_bn_state(self.h, chain, pa, pb, a.size))

    # ---- steps ----
    @staticmethod
    def _idx1(idx0):
        a = np.ascontiguousarray(np.asarray(idx0, dtype=np.int64) + 1)
        return a, a.ctypes.data_as(C.POINTER(C.c_int64))

    def loss_grad(self, idx0):
        a, p = self._idx1(idx0)
        loss = C.c_float(0)
        g = np.empty(self.n_flat, dtype=np.float32)
        self._ck(self.lib.eh_loss_grad(self.h, p, a.size, C.byref(loss), g.ctypes.data_as(_fp)))
        return float(loss.value), g

    def step(self, idx0, want_grad=False):
        a, p = self._idx1(idx0)
        loss = C.c_float(0)
        g = np.empty(self.n_flat, dtype=np.float32) if want_grad else None
        self._ck(self.lib.eh_step(self.h, p, a.size, C.byref(loss), g.ctypes.data_as(_fp) if want_grad else None))
        return (float(loss.value), g) if want_grad else float(loss.value)

    def step_host(self, xf, y):
        X, forc = xf
        Xc, Xp = _f32(X)
        kf, pf = _ptr_array([forc[f] for f in self.model.forcing])
        kt, pt = _ptr_array([y[t] for t in self.model.targets])
        loss = C.c_float(0)
        self._ck(self.lib.eh_step_host(self.h, Xc.shape[0], Xp, pf, pt, C.byref(loss)))
        return float(loss.value)

    def epoch(self, perm0, batchsize, one_based=False):
        """run_epoch! over a host permutation.  ``one_based=True``: ``perm0`` already is what Julia hands over
        (1-based contiguous int64, ideally page-locked) and is passed through without a copy."""
        if one_based:
            a = perm0
            if a.dtype != np.int64 or not a.flags.c_contiguous:
                raise ValueError("one_based permutations must be contiguous int64")
            p = a.ctypes.data_as(C.POINTER(C.c_int64))
        else:
            a, p = self._idx1(perm0)
        nsteps = (a.size + batchsize - 1) // batchsize
        losses = np.empty(nsteps, dtype=np.float32)
        self._ck(self.lib.eh_epoch(self.h, p, a.size, batchsize, losses.ctypes.data_as(_fp)))
        return losses

    def set_perm(self, perm0):
        a, p = self._idx1(perm0)
        self._ck(self.lib.eh_set_perm(self.h, p, a.size))
        self._perm_n = int(a.size)

    def run_steps(self, batchsize, first_step, n_steps):
        losses = np.empty(n_steps, dtype=np.float32)
        self._ck(self.lib.eh_run_steps(self.h, batchsize, first_step, n_steps, losses.ctypes.data_as(_fp)))
        return losses

    # ---- streaming host batches (pinned memory, asynchronous) ----
    def pinned(self, arr):
        """copy ``arr`` into page-locked host memory owned by this session; returns a numpy view"""
        arr = np.ascontiguousarray(arr)
        p = C.c_void_p()
        st = self.lib.eh_host_alloc(C.byref(p), arr.nbytes)
        if st != _abi.EH_OK:
            raise EasyHybridCudaError(st, "eh_host_alloc failed")
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(p)
        buf = (C.c_char * max(arr.nbytes, 1)).from_address(p.value)
        view = np.frombuffer(buf, dtype=arr.dtype, count=arr.size).reshape(arr.shape)
        view[...] = arr
        return view

    def host_batch(self, X, forc, targ):
        """marshal one host batch (X [B,P], [forcing arrays], [target arrays]) once; reusable handle"""
        pf = (_fp * max(len(forc), 1))(*[a.ctypes.data_as(_fp) for a in forc])
        pt = (_fp * max(len(targ), 1))(*[a.ctypes.data_as(_fp) for a in targ])
        return (X.shape[0], X.ctypes.data_as(_fp), pf, pt, (X, forc, targ))

    def step_host_async(self, batch, loss_array, slot):
        """batch: a ``host_batch`` handle (or the raw tuple) -- ideally over views from ``pinned``"""
        if len(batch) == 3:
            batch = self.host_batch(*batch)
        n, Xp, pf, pt, _keep = batch
        lp = C.cast(loss_array.ctypes.data + 4 * slot, _fp)
        st = self.lib.eh_step_host_async(self.h, n, Xp, pf, pt, lp)
        if st != _abi.EH_OK:
            self._ck(st)

    def kernel_variant(self):
        """name of the compiled kernel family serving this model (eh_kernel_variant)"""
        return (self.lib.eh_kernel_variant(self.h) or b"").decode()

    def epoch_variant(self, batch):
        """name of the kernel family serving the persistent launches at this batch size (eh_epoch_variant)"""
        return (self.lib.eh_epoch_variant(self.h, int(batch)) or b"").decode()

    def sync(self):
        self._ck(self.lib.eh_sync(self.h))
        self._inflight = []

    # ---- data parallel (one process per GPU) ----
    def comm_id(self):
        buf = C.create_string_buffer(_abi.EH_COMM_ID_BYTES)
        self._ck(self.lib.eh_comm_id(self.h, buf))
        return bytes(buf.raw)

    def comm_init(self, rank, world, dist=None, ids=None):
        """connect the ranks' inboxes; ``dist`` is an initialised torch.distributed module used only to
        all-gather the IPC handles (any other transport can pass ``ids`` = list of blobs in rank order)"""
        if ids is None:
            mine = self.comm_id()
            ids = [None] * world
            dist.all_gather_object(ids, mine)
        blob = b"".join(ids)
        assert len(blob) == world * _abi.EH_COMM_ID_BYTES
        self._ck(self.lib.eh_comm_init(self.h, rank, world, C.create_string_buffer(blob, len(blob))))
        self.rank, self.world = rank, world

    def dp_exchange_batch_stats(self, batchsize, dist):
        """data-parallel runs with NaN targets, nseLoss or input BatchNorm: make the per-batch data statistics
        global (eh_dp_batch_moments -> sum over ranks -> eh_dp_set_batch_moments).  Call after set_perm."""
        import torch
        nb = (self._perm_n + batchsize - 1) // batchsize
        mom = np.zeros((nb, _abi.EH_DP_MOMENTS), dtype=np.float64)
        self._ck(self.lib.eh_dp_batch_moments(self.h, batchsize, mom.ctypes.data_as(C.POINTER(C.c_double))))
        t = torch.from_numpy(mom)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.all_reduce(t)
        mom = np.ascontiguousarray(t.cpu().numpy())
        self._ck(self.lib.eh_dp_set_batch_moments(self.h, batchsize, mom.ctypes.data_as(C.POINTER(C.c_double))))

    # ---- evaluation ----
    def eval(self, split, want_yhat=True, want_params=False):
        n = self.n[split]
        T = len(self.model.targets)
        yhat = np.empty((T, n), dtype=np.float32) if want_yhat else None
        stats = np.zeros((T, _abi.EH_EVAL_STATS), dtype=np.float64)
        npar = len(self.model.parameters.names)
        par = np.full((npar, n), np.nan, dtype=np.float32) if want_params else None
        self._ck(self.lib.eh_eval(self.h, split, yhat.ctypes.data_as(_fp) if want_yhat else None,
                                  stats.ctypes.data_as(C.POINTER(C.c_double)),
                                  par.ctypes.data_as(_fp) if want_params else None))
        return yhat, stats, par

    # ---- timing ----
    def last_timing(self):
        ms, n, kms = C.c_float(0), C.c_int64(0), C.c_float(0)
        self._ck(self.lib.eh_last_timing(self.h, C.byref(ms), C.byref(n), C.byref(kms)))
        return float(ms.value), int(n.value), float(kms.value)

    def set_profiling(self, on):
        self._ck(self.lib.eh_set_profiling(self.h, int(bool(on))))
