"""CUDA-event timing of each kernel family at the BASELINE sizes (not under a profiler). Prints ms and algorithmic GB/s."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ra_ra_b200 import abi, cuda, ra  # noqa: E402

ex = cuda.CudaExecutor(0)
lib = cuda.lib
dev = ex.device
log2n = int(os.environ.get("LOG2N", "30"))
N = 1 << log2n
m = 1 << (log2n // 2)
n = N // m
a_t = torch.rand(N, device=dev)
b_t = torch.rand(N, device=dev)
a, b = ex.wrap(a_t), ex.wrap(b_t)
A2 = ex.wrap(a_t.view(m, n))
scal = torch.zeros(4, device=dev)
rows_t, cols_t = torch.zeros(n, device=dev), torch.zeros(m, device=dev)
sp = cuda.stream_ptr


def timed(fn, steps=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def red(desc):
    return lambda: cuda.check(lib.racu_reduce_dev(C.byref(desc), abi.RED_SUM, C.c_void_p(scal.data_ptr()), sp()))


def mp(desc):
    return lambda: cuda.check(lib.racu_map(C.byref(desc), sp()))


d_sum, _k1 = ra.flatten(ra.as_expr(a), ex=ex)
d_sqrm, _k2 = ra.flatten(ra.sqrm(a), ex=ex)
d_dot, _k3 = ra.flatten(a * b, ex=ex)
d_rows, _k4 = ra.flatten(ra.as_expr(A2), dst=ex.wrap(rows_t)(ra.insert(1)), assign=abi.ASSIGN_ADD, ex=ex)
d_cols, _k5 = ra.flatten(ra.as_expr(A2), dst=ex.wrap(cols_t), assign=abi.ASSIGN_ADD, ex=ex)
d_copy, _k6 = ra.flatten(ra.as_expr(a), dst=b, assign=abi.ASSIGN_SET, ex=ex)
d_axpy, _k7 = ra.flatten(2.5 * a, dst=b, assign=abi.ASSIGN_ADD, ex=ex)
out = {}
for name, fn, nbytes in (("sum", red(d_sum), 4 * N), ("sqrm", red(d_sqrm), 4 * N), ("dot", red(d_dot), 8 * N),
                         ("sum-rows", mp(d_rows), 4 * N), ("sum-cols", mp(d_cols), 4 * N),
                         ("copy b=a", mp(d_copy), 8 * N), ("axpy b+=2.5*a", mp(d_axpy), 12 * N),
                         ("torch copy_", lambda: b_t.copy_(a_t), 8 * N), ("torch sum", lambda: a_t.sum(), 4 * N),
                         ("torch dot", lambda: torch.dot(a_t, b_t), 8 * N)):
    ms = timed(fn)
    out[name] = {"ms": round(ms, 4), "GB/s": round(nbytes / ms / 1e6, 1), "kernel": lib.racu_last_kernel().decode()}
    print(f"{name:16s} {ms:8.4f} ms {nbytes / ms / 1e6:8.1f} GB/s  {out[name]['kernel']}")
# f64 variants of the folds
a64 = ex.wrap(a_t[: N // 2].double())
b64 = ex.wrap(b_t[: N // 2].double())
n64 = N // 2
d_sum64, _q1 = ra.flatten(ra.as_expr(a64), ex=ex)
d_dot64, _q2 = ra.flatten(a64 * b64, ex=ex)
s64 = torch.zeros(2, device=dev, dtype=torch.float64)
for name, d, nbytes in (("sum f64", d_sum64, 8 * n64), ("dot f64", d_dot64, 16 * n64)):
    ms = timed(lambda: cuda.check(lib.racu_reduce_dev(C.byref(d), abi.RED_SUM, C.c_void_p(s64.data_ptr()), sp())))
    out[name] = {"ms": round(ms, 4), "GB/s": round(nbytes / ms / 1e6, 1), "kernel": lib.racu_last_kernel().decode()}
    print(f"{name:16s} {ms:8.4f} ms {nbytes / ms / 1e6:8.1f} GB/s  {out[name]['kernel']}")
print(json.dumps(out))
