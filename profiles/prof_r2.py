"""Driver for ncu, round 2: a few launches of each kernel at the FULL BASELINE sizes (ncu replays every launch ~40 times; all of these
are 0.2-3 ms). Usage: python profiles/prof_r2.py which ...   which in: dot sum gemv32 gevm32 gemv64 gevm64 rows cols stencil map mapv"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ra_ra_b200 import abi, cuda, ra  # noqa: E402

which = sys.argv[1:]
ex = cuda.CudaExecutor(0)
lib = cuda.lib
dev = ex.device
reps = int(os.environ.get("REPS", "3"))
for w in which:
    if w in ("dot", "sum", "rows", "cols"):
        N = 1 << 30
        a_t = torch.rand(N, device=dev)
        a = ex.wrap(a_t)
        A2 = ex.wrap(a_t.view(1 << 15, 1 << 15))
        if w == "dot":
            b = ex.wrap(torch.rand(N, device=dev))
        o_t = torch.zeros(1 << 15, device=dev)
        for _ in range(reps):
            if w == "dot":
                ra.dot(a, b)
            elif w == "sum":
                ra.sum(a)
            elif w == "cols":
                c = ex.wrap(o_t)
                c += A2
            else:
                r = ex.wrap(o_t)(ra.insert(1))
                r += A2
    elif w in ("gemv32", "gevm32", "gemv64", "gevm64"):
        gdt = torch.float32 if w.endswith("32") else torch.float64
        ng = 16384
        gm = ex.wrap(torch.rand(ng, ng, device=dev, dtype=gdt))
        gx = ex.wrap(torch.rand(ng, device=dev, dtype=gdt))
        gy = ex.wrap(torch.zeros(ng, device=dev, dtype=gdt))
        for _ in range(reps):
            if w.startswith("gemv"):
                ra.gemv(gm, gx, gy)
            else:
                ra.gevm(gx, gm, gy)
    elif w in ("map", "mapv"):
        n = 8192
        A2, B2, C2 = (ex.wrap(torch.rand(n, n, device=dev)) for _ in range(3))
        v = ex.wrap(torch.rand(n, device=dev, dtype=torch.float64))
        for _ in range(reps):
            if w == "map":
                B2 += A2 * C2
            else:
                B2 += A2 * C2 + v
    elif w in ("stiff32", "stiff64"):  # user-written 5-point stencil over shifted slices, odd size: the lane-wise map kernel
        n = 8193
        dt = torch.float32 if w.endswith("32") else torch.float64
        v = ex.wrap(torch.rand(n, n, device=dev, dtype=dt))
        ww = ex.wrap(torch.zeros(n, n, device=dev, dtype=dt))
        i, j = ra.iota(n - 2, 1), ra.iota(n - 2, 1)
        for _ in range(reps):
            ww(i, j).assign(4 * v(i, j) - v(i - 1, j) - v(i, j - 1) - v(i + 1, j) - v(i, j + 1))
    elif w == "stencil":
        n3 = 1024
        bufs = [torch.rand(n3, n3, n3, device=dev), torch.zeros(n3, n3, n3, device=dev)]
        s = abi.StencilDesc()
        s.kind, s.dtype, s.rank = abi.STENCIL_3D7, abi.F32, 3
        for k in range(3):
            s.len[k] = n3
            s.src_stride[k] = s.dst_stride[k] = bufs[0].stride(k)
        for _ in range(reps):
            s.src, s.dst = bufs[0].data_ptr(), bufs[1].data_ptr()
            cuda.check(lib.racu_stencil(C.byref(s), cuda.stream_ptr()))
            bufs.reverse()
    torch.cuda.synchronize()
    print(w, lib.racu_last_kernel().decode())
    torch.cuda.empty_cache()
