"""7-point sweep (bench/bench-stencil3.cc:55-64) on slab shapes of the multi-GPU configs, one GPU, CUDA-event timed:
(1024,1024,1024) the single-GPU config, (130,1024,1024) the 8-GPU slab of 1024^3, (258,2048,2048) the 8-GPU slab of 2048^3.
RACU_ST_BAND / RACU_ST_CI are the kernel's tuning knobs (tiles per band, planes per chunk). 8 B / cell algorithmic."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ra_ra_b200 import abi, cuda  # noqa: E402

lib = cuda.lib
shapes = [tuple(int(x) for x in s.split("x")) for s in os.environ.get("SHAPES", "1024x1024x1024,130x1024x1024,258x2048x2048").split(",")]
out = {}
for shape in shapes:
    A = torch.empty(shape, device="cuda", dtype=torch.float32).uniform_(-1, 1)
    An = torch.zeros_like(A)
    bufs = [A, An]
    sweep = cuda.stencil_sweep()

    def step():
        sweep(bufs[0].data_ptr(), bufs[1].data_ptr(), shape, list(A.stride()), 0, 0)
        bufs.reverse()
    for _ in range(4):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 20
    for _ in range(K):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    cells = shape[0] * shape[1] * shape[2]
    out["x".join(map(str, shape))] = {"ms": round(ms, 4), "Gcells/s": round(cells / ms / 1e6, 1), "GB/s": round(8 * cells / ms / 1e6, 1)}
    del A, An, bufs
print(json.dumps({"band": os.environ.get("RACU_ST_BAND", "default"), "ci": os.environ.get("RACU_ST_CI", "default"), **out}))
