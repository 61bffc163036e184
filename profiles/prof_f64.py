import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ra_ra_b200 import abi, cuda, ra
ex = cuda.CudaExecutor(0)
a = ex.wrap(torch.rand(1 << 27, device=ex.device, dtype=torch.float64))
b = ex.wrap(torch.rand(1 << 27, device=ex.device, dtype=torch.float64))
for _ in range(3):
    ra.sum(a); ra.dot(a, b)
torch.cuda.synchronize()
print("ok")
