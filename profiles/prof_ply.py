"""Small driver for ncu: a few launches of each ply kernel family on one GPU (sizes > L2, short enough for ncu's replays).
Usage: python profiles/prof_ply.py [which ...]   which in: dot sum map mapv mapvf cols rows jit dgemm sgemm stencil"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ra_ra_b200 import abi, cuda, ra  # noqa: E402

which = sys.argv[1:] or ["dot", "sum", "map", "mapv", "cols", "rows", "stencil"]
ex = cuda.CudaExecutor(0)
lib = cuda.lib
dev = ex.device
N = 1 << 28
m = n = 1 << 14
a_t = torch.rand(N, device=dev)
b_t = torch.rand(N, device=dev)
a, b = ex.wrap(a_t), ex.wrap(b_t)
A2, B2 = ex.wrap(a_t.view(m, n)), ex.wrap(b_t.view(m, n))
out = torch.zeros(4, device=dev)
rows_t, cols_t = torch.zeros(n, device=dev), torch.zeros(m, device=dev)
v = ex.wrap(torch.rand(m, device=dev, dtype=torch.float64))
vf = ex.wrap(torch.rand(m, device=dev, dtype=torch.float32))
reps = 3
for w in which:
    for _ in range(reps):
        if w == "dot":
            ra.dot(a, b)
        elif w == "sum":
            ra.sum(a)
        elif w == "map":
            B2 += A2 * A2
        elif w == "mapv":
            B2 += A2 * A2 + v
        elif w == "mapvf":
            B2 += A2 * A2 + vf
        elif w == "cols":
            c = ex.wrap(cols_t)
            c += A2
        elif w == "rows":
            r = ex.wrap(rows_t)(ra.insert(1))
            r += A2
        elif w == "jit":      # a program outside the table: run-time specialised map (where + sqrt)
            B2.assign(ra.where(A2 < B2, A2 * B2, ra.sqrt(A2)) + 0.5)
        elif w == "dgemm":    # FP64 DMMA contraction, 4096^3
            ng = 4096
            ga, gb = (ex.wrap(torch.rand(ng, ng, device=dev, dtype=torch.float64)) for _ in range(2))
            gc = ex.wrap(torch.zeros(ng, ng, device=dev, dtype=torch.float64))
            ra.gemm(ga, gb, gc)
        elif w == "sgemm":    # 3xTF32 contraction, 4096^3
            ng = 4096
            ga, gb = (ex.wrap(torch.rand(ng, ng, device=dev, dtype=torch.float32)) for _ in range(2))
            gc = ex.wrap(torch.zeros(ng, ng, device=dev, dtype=torch.float32))
            ra.gemm(ga, gb, gc)
        elif w in ("gemv32", "gevm32", "gemv64", "gevm64"):   # bench-gemv shapes, 16384^2
            gdt = torch.float32 if w.endswith("32") else torch.float64
            ng = 16384
            gm = ex.wrap(torch.rand(ng, ng, device=dev, dtype=gdt))
            gx = ex.wrap(torch.rand(ng, device=dev, dtype=gdt))
            gy = ex.wrap(torch.zeros(ng, device=dev, dtype=gdt))
            if w.startswith("gemv"):
                ra.gemv(gm, gx, gy)
            else:
                ra.gevm(gx, gm, gy)
            del gm
        elif w == "stencil":
            n3 = 512
            A3 = torch.rand(n3, n3, n3, device=dev)
            An = torch.zeros_like(A3)
            s = abi.StencilDesc()
            s.kind, s.dtype, s.rank = abi.STENCIL_3D7, abi.F32, 3
            for k in range(3):
                s.len[k] = n3
                s.src_stride[k] = s.dst_stride[k] = A3.stride(k)
            s.src, s.dst = A3.data_ptr(), An.data_ptr()
            cuda.check(lib.racu_stencil(C.byref(s), cuda.stream_ptr()))
    torch.cuda.synchronize()
    print(w, lib.racu_last_kernel().decode())
