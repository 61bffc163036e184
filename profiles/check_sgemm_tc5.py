"""f32 contraction on tcgen05 (csrc/gemm_tc5.cu) against an fp64 reference: error maps per 32x32 block for quick diagnosis,
then timing at 4096^3 / 8192^3 / 16384^3 against the mma.sync 3xTF32 kernel and cuBLAS."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ra_ra_b200 import abi, cuda, ra
ex = cuda.CudaExecutor(0); lib = cuda.lib
torch.manual_seed(5)
ok_all = True
if len(sys.argv) > 1 and sys.argv[1] == "mask":
    os.environ["RACU_TC5_MASK"] = "1"
    print("RACU_TC5_MASK=1: hi operands are masked explicitly (default: raw fp32 words, the tensor core drops the low 13 bits itself)")
if len(sys.argv) > 1 and sys.argv[1] == "ss":
    os.environ["RACU_TC5_SS"] = "1"
    print("RACU_TC5_SS=1: a operand from shared memory (SS form)")
for (m, n, k) in [(128, 128, 32), (128, 128, 64), (128, 128, 256), (256, 384, 96), (100, 72, 40), (384, 128, 1000), (1024, 1024, 1024)]:
    a = torch.rand(m, k, device="cuda") * 2 - 1
    b = torch.rand(k, n, device="cuda") * 2 - 1
    c0 = torch.rand(m, n, device="cuda") * 2 - 1
    c = c0.clone()
    try:
        ra.gemm(ex.wrap(a), ex.wrap(b), ex.wrap(c))
        torch.cuda.synchronize()
    except Exception as e:
        print((m, n, k), "FAILED", e); ok_all = False; break
    ref = c0.double() + a.double() @ b.double()
    scale = a.double().abs() @ b.double().abs()
    err = (c.double() - ref).abs()
    bound = (k * 2.0 ** -24 + 2.0 ** -20) * scale + 1e-30
    ratio = float((err / bound).max())
    ok = ratio <= 1.0
    ok_all &= ok
    print((m, n, k), lib.racu_last_kernel().decode(), "max err", float(err.max()), "max err/bound", round(ratio, 3), "OK" if ok else "WRONG")
    if not ok:
        eb = err[: (m // 32) * 32, : (n // 32) * 32].reshape(m // 32, 32, n // 32, 32).amax(dim=(1, 3))
        print("  max err per 32x32 block:\n", eb.cpu().numpy().round(4))
        r0 = err[:8, :8].cpu().numpy(); print("  err[:8,:8]\n", r0.round(4))
        # does it look like a single k-block / hi-only result?
        hi_only = c0.double() + (a.double() @ b.double())
        print("  plain tf32-level error would be ~", float((scale * 2.0 ** -11).max()))
print("ALL OK" if ok_all else "SOME WRONG")
if ok_all:
    def timed(fn, steps=3, warmup=1):
        for _ in range(warmup): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps
    for n in (4096, 8192, 16384):
        a, b = (torch.rand(n, n, device="cuda") * 2 - 1 for _ in range(2))
        c = torch.zeros(n, n, device="cuda")
        A, B, Cc = ex.wrap(a), ex.wrap(b), ex.wrap(c)
        ms = timed(lambda: ra.gemm(A, B, Cc)); k1 = lib.racu_last_kernel().decode()
        os.environ["RACU_NO_TCGEN05"] = "1"
        ms2 = timed(lambda: ra.gemm(A, B, Cc)); k2 = lib.racu_last_kernel().decode()
        del os.environ["RACU_NO_TCGEN05"]
        torch.backends.cuda.matmul.allow_tf32 = False
        ms3 = timed(lambda: torch.matmul(a, b, out=c))
        fl = 2.0 * n ** 3 / 1e9
        print(f"{n}^3: {k1} {ms:.2f} ms {fl / ms:.1f} TFLOP/s | {k2} {ms2:.2f} ms {fl / ms2:.1f} | cuBLAS fp32 {ms3:.2f} ms {fl / ms3:.1f}")
