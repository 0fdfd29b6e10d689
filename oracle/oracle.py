"""ctypes wrapper of the CPU oracle (oracle/eh_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py.  The product package never imports this module.
It reuses the product's descriptor marshalling (easyhybrid_b200._abi / model.build_desc) so that
oracle and CUDA library are driven by the very same eh_model_desc bytes.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libeh_oracle.so")
_lib = None
_fp = C.POINTER(C.c_float)
_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("eh_oracle.c", "eh_oracle_core.inc")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        import easyhybrid_b200  # noqa: F401  (registers the package alias)
        from easyhybrid_b200 import _abi
        L = C.CDLL(_SO)
        L.eho_plan_new.restype = C.c_void_p
        L.eho_plan_new.argtypes = [C.POINTER(_abi.eh_model_desc)]
        L.eho_plan_free.argtypes = [C.c_void_p]
        L.eho_num_params.restype = C.c_int64
        L.eho_num_params.argtypes = [C.c_void_p]
        L.eho_max_threads.restype = C.c_int
        L.eho_loss_grad.restype = C.c_double
        L.eho_loss_grad.argtypes = [C.c_void_p, _fp, C.c_int64, _fp, C.POINTER(_fp), C.POINTER(_fp), _i64p, C.c_int64,
                                    _dp, C.c_int, C.c_int]
        L.eho_train_steps.restype = C.c_int
        L.eho_train_steps.argtypes = [C.c_void_p, _fp, _fp, _fp, _i64p, C.c_int64, _fp, C.POINTER(_fp), C.POINTER(_fp),
                                      _i64p, C.c_int64, C.c_int64, _fp, C.c_int]
        L.eho_opt_step.argtypes = [C.c_void_p, _fp, _fp, _fp, _i64p, _fp]
        L.eho_forward.restype = C.c_int
        L.eho_forward.argtypes = [C.c_void_p, _fp, C.c_int64, _fp, C.POINTER(_fp), _fp, _fp, C.c_int, C.c_int]
        L.eho_get_bn_state.argtypes = [C.c_void_p, _fp, _fp]
        L.eho_set_bn_state.argtypes = [C.c_void_p, _fp, _fp]
        L.eho_loss_fn.restype = C.c_double
        L.eho_loss_fn.argtypes = [C.c_int, _dp, _dp, C.POINTER(C.c_uint8), C.c_int64]
        for name in ("eho_scale_single_param", "eho_scale_single_param_minmax"):
            getattr(L, name).restype = C.c_float
            getattr(L, name).argtypes = [C.c_float] * 3
        L.eho_inv_sigmoid.restype = C.c_float
        L.eho_inv_sigmoid.argtypes = [C.c_float]
        L.eho_hard_sigmoid.restype = C.c_double
        L.eho_hard_sigmoid.argtypes = [C.c_double]
        _lib = L
    return _lib


LOSS_KINDS = {"mse": 0, "rmse": 1, "mae": 2, "nseLoss": 3, "nse": 4, "r2": 5, "pearson": 6, "pearsonLoss": 7,
              "kgeLoss": 8, "kge": 9, "pbkgeLoss": 10, "pbkge": 11, "α": 12, "β": 13}


def loss_fn(yhat, y, mask, kind):
    """reference loss_fn(ŷ, y, y_nan, Val(kind)) in Float64 (src/losses/loss_fn.jl:58-179)."""
    yhat = np.ascontiguousarray(yhat, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    return lib().eho_loss_fn(LOSS_KINDS[kind], yhat.ctypes.data_as(_dp), y.ctypes.data_as(_dp),
                             m.ctypes.data_as(C.POINTER(C.c_uint8)), y.size)


def _ptrs(arrs):
    keep = [np.ascontiguousarray(a, dtype=np.float32) for a in arrs]
    return keep, (_fp * max(len(keep), 1))(*[a.ctypes.data_as(_fp) for a in keep])


class Oracle:
    """CPU restatement of the hybrid training step for one model descriptor."""

    def __init__(self, model, *, training_loss="mse", agg="sum", opt=None, extra_loss=None):
        from easyhybrid_b200.model import build_desc
        self.L = lib()
        self.model = model
        self.bundle = build_desc(model, training_loss=training_loss, agg=agg, opt=opt, extra_loss=extra_loss)
        self.plan = self.L.eho_plan_new(self.bundle.byref())
        if not self.plan:
            raise RuntimeError("oracle: unsupported descriptor")
        self.n_flat = int(self.L.eho_num_params(self.plan))
        self.m = np.zeros(self.n_flat, np.float32)
        self.v = np.zeros(self.n_flat, np.float32)
        self.t = C.c_int64(0)

    def __del__(self):
        if getattr(self, "plan", None):
            self.L.eho_plan_free(self.plan)
            self.plan = None

    def _data(self, xf, y):
        X, forc = xf
        Xc = np.ascontiguousarray(X, dtype=np.float32)
        kf, pf = _ptrs([forc[f] for f in self.model.forcing])
        kt, pt = _ptrs([y[t] for t in self.model.targets]) if y is not None else ([], None)
        return Xc, kf, pf, kt, pt

    def loss_grad(self, flat, xf, y, idx0, precision=32, nthreads=1):
        Xc, kf, pf, kt, pt = self._data(xf, y)
        flat = np.ascontiguousarray(flat, dtype=np.float32)
        idx = np.ascontiguousarray(idx0, dtype=np.int64)
        g = np.empty(self.n_flat, dtype=np.float64)
        L = self.L.eho_loss_grad(self.plan, flat.ctypes.data_as(_fp), Xc.shape[0], Xc.ctypes.data_as(_fp), pf, pt,
                                 idx.ctypes.data_as(_i64p), idx.size, g.ctypes.data_as(_dp), precision, nthreads)
        return L, g

    def train_steps(self, flat, xf, y, perm0, batchsize, nthreads=1):
        """in-place optimiser steps over the permutation; returns per-step losses"""
        Xc, kf, pf, kt, pt = self._data(xf, y)
        assert flat.dtype == np.float32 and flat.flags.c_contiguous
        perm = np.ascontiguousarray(perm0, dtype=np.int64)
        nsteps = (perm.size + batchsize - 1) // batchsize
        losses = np.empty(nsteps, dtype=np.float32)
        rc = self.L.eho_train_steps(self.plan, flat.ctypes.data_as(_fp), self.m.ctypes.data_as(_fp),
                                    self.v.ctypes.data_as(_fp), C.byref(self.t), Xc.shape[0], Xc.ctypes.data_as(_fp),
                                    pf, pt, perm.ctypes.data_as(_i64p), perm.size, batchsize,
                                    losses.ctypes.data_as(_fp), nthreads)
        assert rc == 0
        return losses

    def opt_step(self, flat, grad):
        g = np.ascontiguousarray(grad, dtype=np.float32)
        self.L.eho_opt_step(self.plan, flat.ctypes.data_as(_fp), self.m.ctypes.data_as(_fp), self.v.ctypes.data_as(_fp),
                            C.byref(self.t), g.ctypes.data_as(_fp))

    def forward(self, flat, xf, precision=32, nthreads=1, want_params=False):
        Xc, kf, pf, _, _ = self._data(xf, None)
        n = Xc.shape[0]
        flat = np.ascontiguousarray(flat, dtype=np.float32)
        yhat = np.empty((len(self.model.targets), n), dtype=np.float32)
        par = np.empty((len(self.model.parameters.names), n), dtype=np.float32) if want_params else None
        self.L.eho_forward(self.plan, flat.ctypes.data_as(_fp), n, Xc.ctypes.data_as(_fp), pf,
                           yhat.ctypes.data_as(_fp), par.ctypes.data_as(_fp) if want_params else None, precision, nthreads)
        return (yhat, par) if want_params else yhat

    def bn_state(self):
        n = sum(len(ch["predictors"]) for ch in self.model.chains)
        mean, var = np.empty(n, np.float32), np.empty(n, np.float32)
        self.L.eho_get_bn_state(self.plan, mean.ctypes.data_as(_fp), var.ctypes.data_as(_fp))
        return mean, var


def max_threads():
    return int(lib().eho_max_threads())
