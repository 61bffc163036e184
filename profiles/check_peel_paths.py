"""GPU check of the host-side peel paths of api.cc (more uncollapsible axes than one launch handles) against the CPU oracle:
whole-expression reductions (rewritten as a fold onto a rank-0 destination), `=` onto a broadcast destination (last slice only),
running forms."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle_exec import OracleExecutor
from randexpr import DT, random_view
from ra_ra_b200 import cuda, ra
exc = cuda.CudaExecutor(0)
bad = 0
for seed in range(24):
    shape = [(2, 3, 2, 2, 3, 2), (2,) * 8, (3, 2, 2, 2, 2, 2), (2, 2, 2, 2, 3)][seed % 4]
    res = []
    for ex in (OracleExecutor(), exc):
        rng = np.random.default_rng(4400 + seed)
        k = int(rng.integers(0, len(shape)))
        dst, _ = random_view(rng, ex, shape[:k] + shape[k + 1:], DT[rng.integers(0, 4)], True)
        dst = dst(ra.dots(k), ra.insert(1))
        a, _ = random_view(rng, ex, shape, DT[rng.integers(2, 4)], True)
        b, _ = random_view(rng, ex, shape[:3], np.int32, True)
        form = seed % 4
        if form == 0:
            dst.assign(a * 2 + 1)
        elif form == 1:
            dst += a
        elif form == 2:
            dst.assign(dst + a)
        else:
            dst.assign(ra.max(dst, ra.cast(dst.dtype, a)))
        out = dst.owner.cpu().numpy() if ex.name == "cuda" else np.array(dst.owner, copy=True)
        res.append((out, int(ra.sum(a * b - 3)), int(ra.amax(a * b)), int(ra.amin(a - b))))
    ok = np.array_equal(res[0][0], res[1][0]) and res[0][1:] == res[1][1:]
    bad += not ok
    print(seed, shape, "OK" if ok else ("WRONG", res[0][1:], res[1][1:]), cuda.lib.racu_last_kernel().decode())
print("ALL OK" if not bad else f"{bad} WRONG")
