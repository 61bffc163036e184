"""laplace3d 7-point sweep, f32 n^3 on one B200: T single steps (racu_stencil) against racu_stencil_steps (two steps per pass), Gcells/s per
step, with the bit-exact comparison of both buffers at the same size. RACU_ST2_CI = output planes per CTA of the two-step kernel."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ra_ra_b200 import abi, cuda, ra  # noqa: E402

n = int(os.environ.get("NS", "1024"))
T = int(os.environ.get("TS", "100"))
dt = torch.float64 if os.environ.get("F64") else torch.float32
ex = cuda.CudaExecutor(0)


def desc(x, y):
    s = abi.StencilDesc()
    s.kind, s.dtype, s.rank = abi.STENCIL_3D7, (abi.F32 if dt == torch.float32 else abi.F64), 3
    s.src, s.dst = x.data_ptr(), y.data_ptr()
    for k in range(3):
        s.len[k], s.src_stride[k], s.dst_stride[k] = n, x.stride(k), y.stride(k)
    return s


def singles(a, b, steps):
    for t in range(steps):
        x, y = (b, a) if t & 1 else (a, b)
        cuda.check(cuda.lib.racu_stencil(C.byref(desc(x, y)), cuda.stream_ptr()))


def timed(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


g = torch.Generator(device="cuda")
g.manual_seed(41)
a0 = (torch.rand(n, n, n, device="cuda", generator=g, dtype=dt) * 2 - 1) * 0.125   # (|7-point sweep| <= 12 |a|: scaled so 10 steps stay finite)
out = {"n": n, "T": T, "dtype": str(dt)}
a, b = a0.clone(), torch.zeros_like(a0)
singles(a, b, 10)
ra_, rb_ = a.clone(), b.clone()
a.copy_(a0); b.zero_()
scratch = torch.empty_like(a0)
cuda.stencil_steps(a, b, 10, scratch=scratch)
out["bit_exact_10_steps_both_buffers"] = bool(torch.equal(a, ra_) and torch.equal(b, rb_))
del ra_, rb_
a.copy_(a0 * 0); b.zero_()    # zeros: T = 100 steps of random data overflow; timing does not depend on the values
ms1 = timed(lambda: singles(a, b, T))
out["single_steps"] = {"ms_per_step": round(ms1 / T, 4), "Gcells/s": round(n ** 3 / (ms1 / T) / 1e6, 1)}
for ci in (os.environ.get("CIS", "8").split(",")):
    os.environ["RACU_ST2_CI"] = ci
    ms2 = timed(lambda: cuda.stencil_steps(a, b, T, scratch=scratch))
    out[f"two_per_pass_ci{ci}"] = {"ms_per_step": round(ms2 / T, 4), "Gcells/s": round(n ** 3 / (ms2 / T) / 1e6, 1), "speedup": round(ms1 / ms2, 3)}
    print(ci, out[f"two_per_pass_ci{ci}"], flush=True)
print(json.dumps(out))
