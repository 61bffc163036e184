"""DGEMM forms (RACU_DGEMM = b|s|w + 32x3|16x4|16x3): exactness on the integer-valued inputs of bench/bench-gemm.cc:229-230, then time at NG^3
with cuBLAS (torch.matmul) beside it. One process per form (the knob is read per call, but a fresh process keeps the timings independent)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ra_ra_b200 import cuda, ra  # noqa: E402

n = int(os.environ.get("NG", "8192"))
ex = cuda.CudaExecutor(0)
dev = ex.device


def timed(fn, steps=3, warmup=1):
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
forms = sys.argv[1:] or ["b32x3", "s32x3", "w32x3"]
m = 1024
i0 = torch.arange(m, device=dev, dtype=torch.float64)
a_s = (i0[:, None] - i0[None, :]).contiguous()
b_s = (i0[None, :] - 2 * i0[:, None]).contiguous()
want = a_s @ b_s
a_t = torch.rand(n, n, device=dev, dtype=torch.float64) * 2 - 1
b_t = torch.rand(n, n, device=dev, dtype=torch.float64) * 2 - 1
c_t = torch.zeros(n, n, device=dev, dtype=torch.float64)
ms_cublas = timed(lambda: torch.matmul(a_t, b_t, out=c_t))
ref = c_t.clone()
out["cublas"] = {"ms": round(ms_cublas, 3), "TFLOP/s": round(2.0 * n ** 3 / ms_cublas / 1e9, 2)}
print("cublas", out["cublas"], flush=True)
for f in forms:
    os.environ["RACU_DGEMM"] = f
    c_s = torch.zeros(m, m, device=dev, dtype=torch.float64)
    ra.gemm(ex.wrap(a_s), ex.wrap(b_s), ex.wrap(c_s))
    torch.cuda.synchronize()
    exact = bool(torch.equal(c_s, want))
    c_t.zero_()
    a, b, c = ex.wrap(a_t), ex.wrap(b_t), ex.wrap(c_t)
    ra.gemm(a, b, c)
    torch.cuda.synchronize()
    err = float((c_t - ref).abs().max())
    ms = timed(lambda: ra.gemm(a, b, c), steps=5)
    out[f] = {"ms": round(ms, 3), "TFLOP/s": round(2.0 * n ** 3 / ms / 1e9, 2), "exact_int": exact, "max_abs_diff_vs_cublas": err,
              "kernel": cuda.lib.racu_last_kernel().decode()}
    print(f, out[f], flush=True)
print(json.dumps(out))
