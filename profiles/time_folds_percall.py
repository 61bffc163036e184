"""Per-call timing of small host-returning folds (racu_reduce) and device-resident ones (racu_reduce_dev): latency, not bandwidth."""
import ctypes as C, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ra_ra_b200 import abi, cuda, ra
ex = cuda.CudaExecutor(0); lib = cuda.lib
n = 8192
A, Cc = (ex.wrap(torch.rand(n, n, device="cuda")) for _ in range(2))
out = torch.zeros(4, device="cuda")
import numpy as np
cases = {"amax(abs(A-C))": (ra.abs(A - Cc), abi.RED_AMAX), "sum(where(A>.5,A,0))": (ra.where(A > np.float32(0.5), A, np.float32(0)), abi.RED_SUM),
         "sum(A*C) [table]": (A * Cc, abi.RED_SUM), "amax(A) ": (ra.as_expr(A), abi.RED_AMAX)}
for name, (e, red) in cases.items():
    d, keep = ra.flatten(e, ex=ex)
    host, devt = [], []
    for i in range(12):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ex.reduce(d, red, abi.F32)
        host.append((time.perf_counter() - t0) * 1e3)
    for i in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ex.reduce_dev(d, red, out); e1.record(); torch.cuda.synchronize()
        devt.append(e0.elapsed_time(e1))
    print(f"{name:24s} {lib.racu_last_kernel().decode():18s} host-returning ms: " + " ".join(f"{x:.3f}" for x in host))
    print(f"{'':43s} device-resident ms: " + " ".join(f"{x:.3f}" for x in devt))
