"""configs[2]: strided views on ra::Big<double,3> 1024^3 (transpose, reverse, iota-slice assignment), CUDA-event timed.
Algorithmic bytes: 16 B per touched element. Checks every result against torch indexing of the same data."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ra_ra_b200 import abi, cuda, ra  # noqa: E402

n = int(os.environ.get("N3", "1024"))
ex = cuda.CudaExecutor(0)
lib = cuda.lib
A_t = torch.empty(n, n, n, device=ex.device, dtype=torch.float64)
A = ex.wrap(A_t)
A.assign(ra._0 * (n * n) + ra._1 * n + ra._2)  # A[i,j,k] = i*n^2 + j*n + k: any misplaced element is detectable
B_t = torch.zeros_like(A_t)
B = ex.wrap(B_t)


def timed(fn, steps=5, warmup=2):
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
cases = [
    ("B = A", lambda: B.assign(A), lambda: torch.equal(B_t, A_t), n ** 3),
    ("B = transpose(A,{2,1,0})", lambda: B.assign(ra.transpose(A, (2, 1, 0))), lambda: torch.equal(B_t, A_t.permute(2, 1, 0)), n ** 3),
    ("B = transpose(A,{1,0,2})", lambda: B.assign(ra.transpose(A, (1, 0, 2))), lambda: torch.equal(B_t, A_t.permute(1, 0, 2)), n ** 3),
    ("B = transpose(A,{0,2,1})", lambda: B.assign(ra.transpose(A, (0, 2, 1))), lambda: torch.equal(B_t, A_t.permute(0, 2, 1)), n ** 3),
    ("B = reverse(A,0)", lambda: B.assign(ra.reverse(A, 0)), lambda: torch.equal(B_t, A_t.flip(0)), n ** 3),
    ("B = reverse(A,1)", lambda: B.assign(ra.reverse(A, 1)), lambda: torch.equal(B_t, A_t.flip(1)), n ** 3),
    ("B = reverse(A,2)", lambda: B.assign(ra.reverse(A, 2)), lambda: torch.equal(B_t, A_t.flip(2)), n ** 3),
    ("B(iota(n/2,0,2), all, iota(n,n-1,-1)) = A(iota(n/2), all, all)",
     lambda: B(ra.iota(n // 2, 0, 2), ra.all, ra.iota(n, n - 1, -1)).assign(A(ra.iota(n // 2), ra.all, ra.all)),
     lambda: torch.equal(B_t[0::2].flip(2), A_t[:n // 2]), n ** 3 // 2),
    ("B(all, iota(n/4,1,3), all) = A(all, iota(n/4,2,2), all)   (bench-from pattern)",
     lambda: B(ra.all, ra.iota(n // 4, 1, 3), ra.all).assign(A(ra.all, ra.iota(n // 4, 2, 2), ra.all)),
     lambda: torch.equal(B_t[:, 1:1 + 3 * (n // 4):3], A_t[:, 2:2 + 2 * (n // 4):2]), n ** 3 // 4),
]
for name, fn, check, elems in cases:
    B_t.zero_()
    ms = timed(fn)
    ok = bool(check())
    gbs = 16.0 * elems / ms / 1e6
    out[name] = {"ms": round(ms, 3), "GB/s": round(gbs, 1), "ok": ok, "kernel": lib.racu_last_kernel().decode()}
    print(f"{name:80s} {ms:8.3f} ms {gbs:8.1f} GB/s ok={ok} {out[name]['kernel']}")
print(json.dumps(out))
