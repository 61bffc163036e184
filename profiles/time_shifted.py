"""Slice expressions shifted by one element (user-written stencils: rows not 16-byte aligned, row length not a whole number of
vectors) on the specialised map kernels: the 2-D 7-point mass stencil and the 5-point stiffness stencil of examples/laplace2d.cc at
4097^2 f64, the 1-D 3-point stencil of bench/bench-stencil1.cc at 2^26 f32, a shifted axpy. GB/s = algorithmic bytes (one read of the
source + one write of the destination per element) / time."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ra_ra_b200 import cuda, ra  # noqa: E402

ex = cuda.CudaExecutor(0)
lib = cuda.lib
PEAK = 6543.7


def timed(fn, steps=20, warmup=3):
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


out = {}
for n in [int(x) for x in os.environ.get("NSH", "4097,8193,8196").split(",")]:
    for dt, name, es in ((torch.float64, "f64", 8), (torch.float32, "f32", 4)):
        v = ex.wrap(torch.rand(n, n, device=ex.device, dtype=dt))
        w = ex.wrap(torch.zeros(n, n, device=ex.device, dtype=dt))
        i, j = ra.iota(n - 2, 1), ra.iota(n - 2, 1)
        h = 1.0 / n
        cases = {
            "mass 2D7 (laplace2d.cc:48)": lambda: w(i, j).assign(ra.sqrm(h) * (v(i, j) / 2. + (v(i - 1, j) + v(i, j - 1) + v(i + 1, j) + v(i, j + 1) + v(i + 1, j + 1) + v(i - 1, j - 1)) / 12.)),
            "stiff 2D5 (laplace2d.cc:36)": lambda: w(i, j).assign(4 * v(i, j) - v(i - 1, j) - v(i, j - 1) - v(i + 1, j) - v(i, j + 1)),
            "shifted axpy w(i,j) += 2*v(i,j+1)": lambda: w(i, j).__iadd__(2 * v(i, j + 1)),
        }
        for cname, fn in cases.items():
            ms = timed(fn)
            nb = (3 if "axpy" in cname else 2) * es * (n - 2) ** 2
            key = f"{cname} {name} {n}^2"
            out[key] = {"ms": round(ms, 4), "GB/s": round(nb / ms / 1e6, 1), "frac": round(nb / ms / 1e6 / PEAK, 3), "kernel": lib.racu_last_kernel().decode()}
            print(key, out[key], flush=True)
N = 1 << 26
A = ex.wrap(torch.rand(N, device=ex.device))
An = ex.wrap(torch.zeros(N, device=ex.device))
I = ra.iota(N - 2, 1)
ms = timed(lambda: An(I).assign(-2 * A(I) + A(I + 1) + A(I - 1)))
out["1D3 (bench-stencil1.cc:45) f32 2^26"] = {"ms": round(ms, 4), "GB/s": round(8 * N / ms / 1e6, 1), "frac": round(8 * N / ms / 1e6 / PEAK, 3), "kernel": lib.racu_last_kernel().decode()}
print("1D3", out["1D3 (bench-stencil1.cc:45) f32 2^26"])
Ad = ex.wrap(torch.rand(N, device=ex.device, dtype=torch.float64))
And = ex.wrap(torch.zeros(N, device=ex.device, dtype=torch.float64))
ms = timed(lambda: And(I).assign(-2 * Ad(I) + Ad(I + 1) + Ad(I - 1)))
out["1D3 (bench-stencil1.cc:45) f64 2^26"] = {"ms": round(ms, 4), "GB/s": round(16 * N / ms / 1e6, 1), "frac": round(16 * N / ms / 1e6 / PEAK, 3), "kernel": lib.racu_last_kernel().decode()}
print("1D3 f64", out["1D3 (bench-stencil1.cc:45) f64 2^26"])
print(json.dumps(out))
