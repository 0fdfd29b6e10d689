"""Loader for ``libeasyhybrid_cuda.so``.  There is no fallback: if the CUDA library
is missing or does not export the full ABI, importing the compute path fails loudly."""
from __future__ import annotations

import ctypes
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libeasyhybrid_cuda.so")
_lib = None


class EasyHybridCudaError(RuntimeError):
    def __init__(self, status, message):
        self.status = status
        super().__init__(f"{_abi.STATUS_NAMES.get(status, status)}: {message}")


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C easyhybrid.jl_b200/csrc).  There is no CPU fallback."
            )
        _lib = _abi.declare(ctypes.CDLL(LIB_PATH))
    return _lib
