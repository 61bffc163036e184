"""One GPU, the 8-GPU slab (130 x 1024 x 1024): plain racu_stencil against the FUSED kernel variant (racu_stencil_push with
no neighbours: no pushes, no waits) - isolates what the fused instantiation itself costs, and the host launch rate."""
import ctypes as C, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ra_ra_b200 import abi, cuda
lib = cuda.lib
shape = tuple(int(x) for x in os.environ.get("SHAPE", "130x1024x1024").split("x"))
A = torch.empty(shape, device="cuda").uniform_(-1, 1); B = torch.zeros_like(A)
flags = torch.zeros(8, device="cuda", dtype=torch.int32)
def desc(src, dst):
    s = abi.StencilDesc(); s.kind, s.dtype, s.rank, s.src, s.dst = abi.STENCIL_3D7, abi.F32, 3, src.data_ptr(), dst.data_ptr()
    for k in range(3): s.len[k], s.src_stride[k], s.dst_stride[k] = shape[k], src.stride(k), src.stride(k)
    return s
ds = [desc(A, B), desc(B, A)]
h = abi.HaloPush(); h.flags = flags.data_ptr()
sp = cuda.stream_ptr()
def plain(t): cuda.check(lib.racu_stencil(C.byref(ds[t & 1]), sp))
def fused(t):
    h.step = t
    cuda.check(lib.racu_stencil_push(C.byref(ds[t & 1]), C.byref(h), sp))
for name, fn in (("plain", plain), ("fused(no neighbours)", fused), ("plain", plain), ("fused(no neighbours)", fused)):
    for t in range(10): fn(t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 400
    t0 = time.perf_counter(); e0.record()
    for t in range(K): fn(t)
    t_issue = time.perf_counter() - t0
    e1.record(); torch.cuda.synchronize()
    print(f"{name:22s} {e0.elapsed_time(e1) / K:.4f} ms/step on the device, host issue {1e3 * t_issue / K:.4f} ms/step, kernel {lib.racu_last_kernel().decode()}")
