"""Interpreter-path timings (programs without a static specialisation): mixed types, where/pick, long expressions."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ra_ra_b200 import abi, cuda, ra
ex = cuda.CudaExecutor(0); lib = cuda.lib; dev = ex.device
n = 8192
A, B, Cc = (ex.wrap(torch.rand(n, n, device=dev)) for _ in range(3))
v64 = ex.wrap(torch.rand(n, device=dev, dtype=torch.float64))
def timed(fn, steps=10, warmup=3):
    for _ in range(warmup): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
def run(name, fn, nbytes):
    ms = timed(fn)
    print(f"{name:58s} {ms:8.4f} ms {nbytes / ms / 1e6:8.1f} GB/s  {lib.racu_last_kernel().decode()}")
E = n * n
run("B += A*C + v(double)", lambda: B.__iadd__(A * Cc + v64), 16 * E)
run("B += 2.5*A (double scalar)", lambda: B.__iadd__(2.5 * A), 12 * E)
run("B = where(A > C, A, C)", lambda: B.assign(ra.where(A > Cc, A, Cc)), 12 * E)
run("B = sqrt(A*A + C*C)", lambda: B.assign(ra.sqrt(A * A + Cc * Cc)), 12 * E)
run("B = (A - C)*(A + C) + A*0.5f - C", lambda: B.assign((A - Cc) * (A + Cc) + A * ra.np.float32(0.5) - Cc), 12 * E)
run("sum(where(A > 0.5f, A, 0f))", lambda: ra.sum(ra.where(A > ra.np.float32(0.5), A, ra.np.float32(0))), 4 * E)
run("B = A(reversed rows) * C  [int idx]", lambda: B.assign(ra.reverse(A, 0) * Cc + ra._1), 12 * E)
I32 = ex.wrap(torch.randint(-5, 6, (n, n), device=dev, dtype=torch.int32))
D = ex.wrap(torch.rand(n // 2, n, device=dev, dtype=torch.float64))
Ah = ex.wrap(A.owner[: n // 2])
run("B = pick(int(A>.5)+int(C>.5), A, C, A+C)", lambda: B.assign(ra.pick(ra.cast(abi.I32, A > 0.5) + ra.cast(abi.I32, Cc > 0.5), A, Cc, A + Cc)), 12 * E)
run("I = (I*3 - I%4) & 7", lambda: I32.assign((I32 * 3 - I32 % 4) & 7), 8 * E)
run("D(f64) = A(f32)*D + 1", lambda: D.assign(Ah * D + 1.0), (4 + 8 + 8) * E // 2)
run("amax(abs(A - C))", lambda: ra.amax(ra.abs(A - Cc)), 8 * E)
# unbeaten subscripts (gather): whole rows by index (bandwidth bound), random flat permutation (32-byte sector per 4-byte element)
ridx = ex.wrap(torch.randint(0, n, (n,), device=dev, dtype=torch.int32))
perm = ex.wrap(torch.randperm(n * n, device=dev))
flatA, flatB = ex.wrap(A.owner.view(-1)), ex.wrap(B.owner.view(-1))
run("B = A(ridx)   [rows by index]", lambda: B.assign(A(ridx)), 8 * E)
run("b = a(perm)   [random permutation]", lambda: flatB.assign(flatA(perm)), 16 * E)
