"""Row / column folds at the BASELINE sizes on one B200: gemv, gevm (ra/ra.hh:414-438) f32 / f64 at 16384^2 and 8192^2, bench-sum-rows /
bench-sum-cols f32 at 32768^2; sweeps the planner's tuning knobs (read per call). HBM-bound: GB/s = matrix bytes / time."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ra_ra_b200 import abi, cuda, ra  # noqa: E402

os.environ.setdefault("RACU_NO_PLAN_CACHE", "1")  # the knobs below are read at plan time
ex = cuda.CudaExecutor(0)
lib = cuda.lib
PEAK = 6543.7
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timed(fn, steps=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def knobs(**kw):
    for k in ("RACU_ROWS_MINROWS", "RACU_ROWS_MAXVEC", "RACU_OUTER_CTAS"):
        os.environ.pop(k, None)
    for k, v in kw.items():
        os.environ[k] = str(v)


CONFIGS = [("default", {}), ("rows off", {"RACU_ROWS_MINROWS": 1 << 40}), ("rows maxvec 2048", {"RACU_ROWS_MAXVEC": 2048}),
           ("outer ctas 888", {"RACU_OUTER_CTAS": 888}), ("outer ctas 1184", {"RACU_OUTER_CTAS": 1184})]
only = os.environ.get("ONLY")
out = {}
for n in (16384, 8192):
    for dt, name, es in ((torch.float32, "f32", 4), (torch.float64, "f64", 8)):
        a_t = torch.rand(n, n, device=ex.device, dtype=dt) * 2 - 1
        a = ex.wrap(a_t)
        x = ex.wrap(torch.rand(n, device=ex.device, dtype=dt))
        y = ex.wrap(torch.zeros(n, device=ex.device, dtype=dt))
        for nm, fn in (("gemv", lambda: ra.gemv(a, x, y)), ("gevm", lambda: ra.gevm(x, a, y))):
            for cname, kw in CONFIGS:
                if only and only != cname:
                    continue
                if ("rows" in cname and nm == "gevm") or ("outer" in cname and nm == "gemv"):
                    continue
                knobs(**kw)
                ms = timed(fn)
                gbs = es * n * n / ms / 1e6
                key = f"{nm} {name} {n}^2 [{cname}]"
                out[key] = {"ms": round(ms, 4), "GB/s": round(gbs, 1), "frac": round(gbs / PEAK, 3), "kernel": lib.racu_last_kernel().decode()}
                print(key, out[key], flush=True)
        del a_t, a
        torch.cuda.empty_cache()
n = 32768
a_t = torch.rand(n, n, device=ex.device, dtype=torch.float32)
A = ex.wrap(a_t)
rows = ex.wrap(torch.zeros(n, device=ex.device, dtype=torch.float32))
cols = ex.wrap(torch.zeros(n, device=ex.device, dtype=torch.float32))
d_rows, k1 = ra.flatten(ra.as_expr(A), dst=rows(ra.insert(1)), assign=abi.ASSIGN_ADD, ex=ex)
d_cols, k2 = ra.flatten(ra.as_expr(A), dst=cols, assign=abi.ASSIGN_ADD, ex=ex)
import ctypes as C  # noqa: E402
for nm, d in (("sum-rows", d_rows), ("sum-cols", d_cols)):
    for cname, kw in CONFIGS:
        if ("rows" in cname and nm == "sum-rows") or ("outer" in cname and nm == "sum-cols"):
            continue
        knobs(**kw)
        ms = timed(lambda: cuda.check(lib.racu_map(C.byref(d), cuda.stream_ptr())))
        gbs = 4 * n * n / ms / 1e6
        key = f"{nm} f32 {n}^2 [{cname}]"
        out[key] = {"ms": round(ms, 4), "GB/s": round(gbs, 1), "frac": round(gbs / PEAK, 3), "kernel": lib.racu_last_kernel().decode()}
        print(key, out[key], flush=True)
knobs()
print(json.dumps(out))
