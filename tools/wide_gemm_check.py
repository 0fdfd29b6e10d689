"""GPU self-test of the tcgen05 GEMM kernels of the wide path (eh_selftest_wide_gemm) against torch."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(ROOT, "easyhybrid.jl_b200", "libeasyhybrid_cuda.so"))
lib.eh_selftest_wide_gemm.restype = C.c_int
lib.eh_selftest_wide_gemm.argtypes = [C.c_int32] * 6 + [C.c_void_p] * 5 + [C.c_int32, C.c_void_p]


def bf(x):
    return x.to(torch.bfloat16)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def run(mode, M, N, K, ks=1, act=1, time_it=False):
    g = torch.Generator().manual_seed(1234 + mode % 3)
    base = mode - 3 if mode >= 3 else mode
    if base < 2:
        A = bf(torch.randn(M, K, generator=g) * 0.5)
        B = bf(torch.randn(N, K, generator=g) * (1.0 / K ** 0.5))
    else:
        A = bf(torch.randn(K, M, generator=g) * 0.5)      # deltas  [batch x M]
        B = bf(torch.randn(K, N, generator=g) * 0.5)      # acts    [batch x N]
    bias = torch.randn(N, generator=g) * 0.1
    aux = bf(torch.tanh(torch.randn(M, N, generator=g)))
    ms = C.c_float(0)
    if base == 2:
        out = torch.empty(ks, M, N, dtype=torch.float32)
    else:
        out = torch.empty(M, N, dtype=torch.bfloat16)
    st = lib.eh_selftest_wide_gemm(mode, M, N, K, ks, act, ptr(A), ptr(B), ptr(bias), ptr(aux), ptr(out), 0,
                                   C.byref(ms) if time_it else None)
    assert st == 0, f"status {st}"
    dev = "cuda"
    Af, Bf = A.float().to(dev), B.float().to(dev)
    if base == 0:
        z = Af @ Bf.T + bias.to(dev)
        ref = torch.tanh(z) if act == 1 else (torch.sigmoid(z) if act == 2 else z)
        got = out.float().to(dev)
        tol = 1e-2
    elif base == 1:
        a = aux.float().to(dev)
        ref = (Af @ Bf.T) * (1 - a * a)
        got = out.float().to(dev)
        tol = 1e-2
    else:
        ref = Af.T @ Bf
        got = out.to(dev).sum(0)
        tol = 2e-3
    err = (got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-9)
    flops = 2.0 * M * N * K
    msg = f"mode {mode} M {M} N {N} K {K} ks {ks}: max rel err {err:.3e}"
    if time_it:
        msg += f"  {ms.value * 1e3:.1f} us  {flops / (ms.value * 1e-3) / 1e12:.1f} TFLOP/s"
    print(msg, flush=True)
    return err < tol


def main(perf=False):
    ok = True
    ok &= run(0, 256, 256, 128)
    ok &= run(0, 512, 512, 512)
    ok &= run(0, 384, 256, 256, act=2)
    ok &= run(1, 512, 512, 512)
    ok &= run(3, 256, 256, 128)
    ok &= run(3, 2048, 512, 512)
    ok &= run(4, 2048, 512, 512)
    ok &= run(2, 256, 256, 256, ks=1)
    ok &= run(2, 512, 512, 4096, ks=4)
    if perf:
        run(0, 65536, 512, 512, time_it=True)
        run(1, 65536, 512, 512, time_it=True)
        run(2, 512, 512, 65536, ks=16, time_it=True)
        run(3, 65536, 512, 512, time_it=True)
        run(4, 65536, 512, 512, time_it=True)
    print("WIDE GEMM", "OK" if ok else "FAILED")
    return ok


if __name__ == "__main__":
    sys.exit(0 if main(len(sys.argv) > 1) else 1)
