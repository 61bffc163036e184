"""examples/laplace3d.cc:175-176 on the device: iterations and error of the FEM CG at n = 50, five runs (the scatter-add combines repeated
indices with atomics: does the order ever cost an iteration?)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import kat  # noqa: E402
from ra_ra_b200 import cuda  # noqa: E402

ex = cuda.CudaExecutor(0)
for k in range(5):
    err, its = kat._laplace3d_fem(ex, 50)
    print("run", k, "err", err, "its", its, flush=True)
