"""configs[4]: dense contraction f64 / f32 n^2 on one B200; torch.matmul (cuBLAS) as the practical ceiling."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ra_ra_b200 import abi, cuda, ra  # noqa: E402

n = int(os.environ.get("NG", "16384"))
ex = cuda.CudaExecutor(0)
lib = cuda.lib


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
for dt, name in ((torch.float64, "f64"), (torch.float32, "f32")):
    a_t = torch.rand(n, n, device=ex.device, dtype=dt) * 2 - 1
    b_t = torch.rand(n, n, device=ex.device, dtype=dt) * 2 - 1
    c_t = torch.zeros(n, n, device=ex.device, dtype=dt)
    a, b, c = ex.wrap(a_t), ex.wrap(b_t), ex.wrap(c_t)
    if os.environ.get('SKIP_GEMM'):
        c_t = None
    ms = timed(lambda: ra.gemm(a, b, c)) if not os.environ.get('SKIP_GEMM') else 0.0
    kern = lib.racu_last_kernel().decode()
    flops = 2.0 * n ** 3
    torch.backends.cuda.matmul.allow_tf32 = False
    ms_cublas = timed(lambda: torch.matmul(a_t, b_t, out=c_t)) if not os.environ.get('SKIP_GEMM') else 1.0
    row = {"ms": round(ms, 2), "TFLOP/s": round(flops / max(ms, 1e-9) / 1e9, 2), "kernel": kern, "cublas_ms": round(ms_cublas, 2),
           "cublas_TFLOP/s": round(flops / ms_cublas / 1e9, 2)}
    if dt == torch.float32:
        torch.backends.cuda.matmul.allow_tf32 = True
        ms_tf32 = timed(lambda: torch.matmul(a_t, b_t, out=c_t)) if not os.environ.get('SKIP_GEMM') else 1.0
        row["cublas_tf32_TFLOP/s"] = round(flops / ms_tf32 / 1e9, 2)
    out[f"gemm {name} {n}^2"] = row
    print(f"gemm {name} {n}^2", row)
    # gemv / gevm: HBM bound
    x = ex.wrap(torch.rand(n, device=ex.device, dtype=dt))
    y = ex.wrap(torch.zeros(n, device=ex.device, dtype=dt))
    for nm, fn in (("gemv", lambda: ra.gemv(a, x, y)), ("gevm", lambda: ra.gevm(x, a, y))):
        ms = timed(fn, steps=20, warmup=3)
        es = 8 if dt == torch.float64 else 4
        row = {"ms": round(ms, 3), "GB/s": round(es * n * n / ms / 1e6, 1), "kernel": lib.racu_last_kernel().decode()}
        out[f"{nm} {name} {n}^2"] = row
        print(f"{nm} {name} {n}^2", row)
    del a_t, b_t, c_t
print(json.dumps(out))
