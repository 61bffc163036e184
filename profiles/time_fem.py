"""The FEM operator of examples/laplace3d.cc:60-68 (gather -> 8x8 element matrix -> scatter-add) at n^3 elements, f64: time per application
and GB/s over the bytes its four statements must move (T = 0; T += M . v(EV); w = 0; w(EV) += T; w(bnd) = 0; w *= h).
RACU_FEM_N = elements per edge."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ra_ra_b200 import cuda, ra  # noqa: E402

n = int(os.environ.get("RACU_FEM_N", "128"))
ex = cuda.CudaExecutor(0)
lib = cuda.lib
ISF = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]])
pos = lambda m, i, j, k: m * (m * i + j) + k  # noqa: E731
ei, ej, ek = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
EV = np.stack([pos(n + 1, ei + ISF[v][0], ej + ISF[v][1], ek + ISF[v][2]).ravel() for v in range(8)], 1).astype(np.int32)
nv, ne = (n + 1) ** 3, n ** 3
rng = np.random.default_rng(1)
M = rng.uniform(-1, 1, (8, 8))
ev, Md = ex.from_numpy(EV), ex.from_numpy(M)
v, w = ex.from_numpy(rng.uniform(-1, 1, nv)), ex.from_numpy(np.zeros(nv))
T = ex.from_numpy(np.zeros((ne, 8)))
kern = {}


def mult():
    T.assign(0.)
    t = T
    t += Md(ra.insert(1)) * v(ev(ra.all, ra.insert(1)))
    kern["fold"] = lib.racu_last_kernel().decode()
    w.assign(0.)
    g = w(ev)
    g += T
    kern["scatter"] = lib.racu_last_kernel().decode()
    w.__imul__(0.5)


for _ in range(3):
    mult()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    mult()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
# T = 0 (64 ne W), fold (read EV 32 ne, T read + write 128 ne, gathered v: 8 nv compulsory), w = 0 (8 nv), scatter (EV 32 ne, T 64 ne, w RMW 16 nv), scale (16 nv)
nbytes = 64 * ne + (32 + 128) * ne + 8 * nv + 8 * nv + (32 + 64) * ne + 16 * nv + 16 * nv
out = {"n": n, "elements": ne, "ms_per_application": round(ms, 4), "GB/s_over_compulsory_bytes": round(nbytes / ms / 1e6, 1), "kernels": kern}
print(json.dumps(out))
