#!/usr/bin/env python
"""bench.py - ra-ra traversal hot path on B200: `ply map/reduce HBM GB/s (% of peak); laplace3d Gcells/s`.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload reduce|stencil]

One "step" is one pass of the hot path over one batch of synthetic input. Default workload (BASELINE.json configs[1]):
reductions on 2^30 floats per GPU - ra::sum, reduce_sqrm, dot (ra/ra.hh:348-401) plus bench-sum-rows / bench-sum-cols on
the same buffer seen as 32768x32768 (bench/bench-sum-rows.cc, bench/bench-sum-cols.cc). Algorithmic bytes per step and
GPU: 4 + 4 + 8 + 4 + 4 GiB. N>1: one process per GPU (torchrun), leading axis sharded, per-GPU work fixed (weak scaling),
partials combined with ONE NCCL allreduce through the library's communicator. Inputs (>= 4 GiB per array) are far larger
than L2 (126 MB), so no flush is needed between iterations.

Prints ONE JSON line (rank 0):
  value        whole-job GB/s with the inputs resident in HBM
  e2e          the same metric through the public API with pinned HOST inputs (H2D inside the timed region, results read back)
  e2e_cxx      the same workload through the C++ host header (include/ra/cuda.hh) from host std::vector inputs (N=1)
  roofline     the dominant kernel (sflat_fold_inner<f32, dot> + ply_finish) timed with CUDA events inside the timed steps
  cpu_baseline the reference's loops (oracle/libbaseline.so, single thread like the reference) on a bounded sample (N=1)
  laplace3d    the other half of BASELINE.json's metric at this N: 7-point sweep on a 1024^3 global grid (strong scaling), with a
               `check` that the sharded result equals the single-GPU result bit for bit; at N=8 also 2048^3 (`laplace3d_weak`)
  map_sharded  (N>1) configs[0] fused map, rows sharded, no collective
  extra        (N=1) all five configs on one GPU, each with its kernel, roofline fraction and the CPU loop timed beside it
`--impl reference` times the reference's CPU loops (the port: the reference itself needs gcc >= 14) on the SAME workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GIB = 1 << 30
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch (ncu --set full), see profiles/
DOT_TRAFFIC_BYTES = 8_590_148_000 + 5_600_000   # sflat_fold_inner<f32, dot> at 2^30 elements: profiles/r01_dot_full_summary.md
STENCIL_TRAFFIC = {}                            # filled from profiles/r02_stencil_1024_full_summary.md when captured (see load_traffic)


def load_traffic():
    """Per-launch DRAM traffic of the kernels the roofline objects name, from the committed ncu summaries (profiles/*.json)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    except Exception:
        return {}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="reduce", choices=["reduce", "stencil"])
    ap.add_argument("--log2n", type=int, default=30, help="elements per GPU for the reduce workload")
    ap.add_argument("--extras", type=int, default=1, help="also time the other configs (reported under 'extra')")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-log2n", type=int, default=0, help="CPU sample size; 0 = the workload's own size for --impl reference, 2^28 for the cpu_baseline leg")
    ap.add_argument("--n3", type=int, default=1024, help="global grid edge for the laplace3d leg / --workload stencil")
    ap.add_argument("--n0", type=int, default=0, help="stencil: extent of the sharded axis if not n3 (e.g. 512 at N=4 gives every GPU the slab it has at N=8 on 1024^3)")
    ap.add_argument("--no-overlap", action="store_true", help="stencil: exchange then sweep, no stream overlap")
    ap.add_argument("--laplace", type=int, default=1, help="reduce workload: also run the laplace3d 1024^3 strong-scaling leg (reported under 'laplace3d')")
    ap.add_argument("--gemm-n", type=int, default=16384, help="extras: edge of the gemm / gemv matrices (BASELINE configs[4]: 16384)")
    ap.add_argument("--halo", default="push", choices=["push", "nccl"],
                    help="stencil at N>1: 'push' = one kernel per step, the sweep stores its edge planes into the neighbours' ghosts over "
                         "NVLink peer memory (racu_stencil_push); 'nccl' = ncclSend/Recv exchange overlapped with the interior sweep")
    return ap.parse_args()


# ---------------------------------------------------------------- CPU baseline (reference arm)

F32P, F64P = C.POINTER(C.c_float), C.POINTER(C.c_double)


def _bind_baseline(L):
    L.base_sum_f32.restype = C.c_float
    L.base_sum_f32.argtypes = [F32P, C.c_int64]
    L.base_sqrm_f32.restype = C.c_float
    L.base_sqrm_f32.argtypes = [F32P, C.c_int64]
    L.base_dot_f32.restype = C.c_float
    L.base_dot_f32.argtypes = [F32P, F32P, C.c_int64]
    L.base_sum_cols_f32.argtypes = [F32P, F32P, C.c_int64, C.c_int64]
    L.base_sum_rows_f32.argtypes = [F32P, F32P, C.c_int64, C.c_int64]
    L.base_map_f32.argtypes = [F32P, F32P, F32P, F64P, C.c_int64, C.c_int64]
    L.base_stencil3_f32.argtypes = [F32P, F32P, C.c_int64, C.c_int64, C.c_int64]
    L.base_transpose210_f64.argtypes = [F64P, F64P, C.c_int64]
    L.base_gemm_f64.argtypes = [F64P, F64P, F64P, C.c_int64, C.c_int64, C.c_int64]
    L.base_gemv_f64.argtypes = [F64P, F64P, F64P, C.c_int64, C.c_int64]
    return L


def baseline_lib():
    path = os.path.join(ROOT, "oracle", "libbaseline.so")
    if not os.path.exists(path):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    return _bind_baseline(C.CDLL(path))


def baseline_lib_native():
    """The same loops built -O3 -march=native ON THIS HOST (what the reference's SCons build adds, test/SConstruct:42-48): compiled
    here because the container the repository is built in has another CPU. None if there is no compiler."""
    out = os.path.join(ROOT, "oracle", "_native", "libbaseline_native.so")
    try:
        if not os.path.exists(out):
            os.makedirs(os.path.dirname(out), exist_ok=True)
            subprocess.run([os.environ.get("CXX", "g++"), "-O3", "-march=native", "-std=c++20", "-fPIC", "-shared", "-o", out,
                            os.path.join(ROOT, "oracle", "baseline_loops.cc")], check=True, capture_output=True, timeout=120)
        return _bind_baseline(C.CDLL(out))
    except Exception:
        return None


def cpu_reduce_step(L, a, b, m, n, r, c):
    """One pass of the reduce workload on the reference's loops. Returns algorithmic bytes."""
    pa, pb = a.ctypes.data_as(F32P), b.ctypes.data_as(F32P)
    N = a.size
    L.base_sum_f32(pa, N)
    L.base_sqrm_f32(pa, N)
    L.base_dot_f32(pa, pb, N)
    L.base_sum_rows_f32(r.ctypes.data_as(F32P), pa, m, n)
    L.base_sum_cols_f32(c.ctypes.data_as(F32P), pa, m, n)
    return 4 * N * 6


def _median(ts):
    return sorted(ts)[len(ts) // 2]


def _cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_baseline(log2n, runs=3):
    """cpu_baseline leg of the GPU arm: the reduce workload on the reference's loops, bounded sample, single thread; and the same
    loops built -march=native on this host."""
    L = baseline_lib()
    N = 1 << log2n
    m = 1 << (log2n // 2)
    n = N // m
    rng = np.random.default_rng(0x5EED0011)
    a = rng.random(N, dtype=np.float32)
    b = rng.random(N, dtype=np.float32)
    r, c = np.zeros(n, np.float32), np.zeros(m, np.float32)

    def run(lib):
        ts = []
        for _ in range(runs):
            t0 = time.perf_counter()
            nbytes = cpu_reduce_step(lib, a, b, m, n, r, c)
            ts.append(time.perf_counter() - t0)
        t = _median(ts)  # median of runs, like Benchmark().runs(3) (ra/test.hh:297-307)
        return nbytes / t / 1e9, t
    v, t = run(L)
    out = {"value": v, "unit": "GB/s", "cores": 1, "kind": "port",
           "sample": f"sum+sqrm+dot+sum-rows+sum-cols on 2^{log2n} f32 ({m}x{n}) of the 2^30 workload, median of {runs} runs, "
                     f"oracle/libbaseline.so -O3, single thread (the reference is single threaded; it needs gcc>=14 "
                     f"and cannot be compiled in this image)",
           "seconds": t, "host_cpu": _cpu_model(), "nproc": os.cpu_count()}
    Ln = baseline_lib_native()
    if Ln is not None:
        vn, tn = run(Ln)
        out["march_native"] = {"value": vn, "unit": "GB/s", "cores": 1, "flags": "-O3 -march=native, compiled on this host", "seconds": tn}
    return out


def cpu_all_cores(log2n=24, max_threads=64):
    """Not the reference (which is single threaded, README.md:81): the same loops on disjoint shards, one per host core,
    to show what the box's cores deliver together when a user shards by hand. ctypes releases the GIL during the calls."""
    L = baseline_lib()
    T = max(1, min(os.cpu_count() or 1, max_threads))
    N = 1 << log2n
    m = 1 << (log2n // 2)
    n = N // m
    rng = np.random.default_rng(0x5EED0013)
    shards = [(rng.random(N, dtype=np.float32), rng.random(N, dtype=np.float32), np.zeros(n, np.float32), np.zeros(m, np.float32))
              for _ in range(T)]
    reps = 4

    def work(sh):
        for _ in range(reps):
            cpu_reduce_step(L, sh[0], sh[1], m, n, sh[2], sh[3])
    ths = [threading.Thread(target=work, args=(sh,)) for sh in shards]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    return {"value": T * reps * 4 * N * 6 / dt / 1e9, "unit": "GB/s", "cores": T,
            "sample": f"{T} threads x {reps} passes over private 2^{log2n}-element shards (not the reference: it is single threaded)"}


def cpu_leg_baselines(gemm_n):
    """The reference's loop for every OTHER config, timed on bounded samples (single thread, -O3), to sit beside the GPU numbers of
    `extra`. Each entry says what the sample was; gemm is timed at 1024^3 and the 16384^3 time extrapolated (stated)."""
    L = baseline_lib()
    rng = np.random.default_rng(0x5EED0099)
    out = {}

    def best(fn, runs=3):
        ts = []
        for _ in range(runs):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return _median(ts)
    n = 4096  # map: 16 B / element
    A, B, Cc = (rng.random((n, n), dtype=np.float32) for _ in range(3))
    v = rng.random(n)
    t = best(lambda: L.base_map_f32(B.ctypes.data_as(F32P), A.ctypes.data_as(F32P), Cc.ctypes.data_as(F32P), v.ctypes.data_as(F64P), n, n))
    out["map B+=A*C+v(double)"] = {"GB/s": 16.0 * n * n / t / 1e9, "sample": f"f32 {n}^2 of 8192^2, median of 3", "seconds": t}
    n3 = 384  # stencil: 8 B / cell
    a = rng.random((n3, n3, n3), dtype=np.float32)
    an = np.zeros_like(a)
    t = best(lambda: L.base_stencil3_f32(an.ctypes.data_as(F32P), a.ctypes.data_as(F32P), n3, n3, n3))
    out["stencil3d7"] = {"Gcells/s": float(n3) ** 3 / t / 1e9, "GB/s_algorithmic": 8.0 * n3 ** 3 / t / 1e9, "sample": f"f32 {n3}^3 of 1024^3, median of 3", "seconds": t}
    del a, an
    nt = 384  # transpose {2,1,0}: 16 B / element, source read with stride n^2
    a = rng.random((nt, nt, nt))
    b = np.zeros_like(a)
    t = best(lambda: L.base_transpose210_f64(b.ctypes.data_as(F64P), a.ctypes.data_as(F64P), nt))
    out["views B=transpose(A,{2,1,0})"] = {"GB/s": 16.0 * nt ** 3 / t / 1e9, "sample": f"f64 {nt}^3 of 1024^3, median of 3", "seconds": t}
    del a, b
    ng = 1024  # gemm: i, k, j loops, c row re-streamed K times
    a, b = rng.random((ng, ng)), rng.random((ng, ng))
    c = np.zeros((ng, ng))
    t = best(lambda: L.base_gemm_f64(c.ctypes.data_as(F64P), a.ctypes.data_as(F64P), b.ctypes.data_as(F64P), ng, ng, ng), runs=1)
    out["gemm f64"] = {"GFLOP/s": 2.0 * ng ** 3 / t / 1e9, "sample": f"f64 {ng}^3, one run", "seconds": t,
                       f"extrapolated_seconds_at_{gemm_n}^3": t * (gemm_n / ng) ** 3, "extrapolation": "cubic in n at the measured rate (optimistic: the 16384 rows no longer fit any cache)"}
    nv = 8192
    a, x = rng.random((nv, nv)), rng.random(nv)
    y = np.zeros(nv)
    t = best(lambda: L.base_gemv_f64(y.ctypes.data_as(F64P), a.ctypes.data_as(F64P), x.ctypes.data_as(F64P), nv, nv))
    out["gemv f64"] = {"GB/s": 8.0 * nv * nv / t / 1e9, "sample": f"f64 {nv}^2 of {gemm_n}^2, median of 3", "seconds": t}
    return out


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (port) on the host cores, on the workload of our arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    L = baseline_lib()
    log2n = args.cpu_log2n or args.log2n
    N = 1 << log2n
    m = 1 << (log2n // 2)
    n = N // m
    rng = np.random.default_rng(0x5EED0011)
    a = rng.random(N, dtype=np.float32)
    b = rng.random(N, dtype=np.float32)
    r, c = np.zeros(n, np.float32), np.zeros(m, np.float32)
    for _ in range(args.warmup):
        cpu_reduce_step(L, a, b, m, n, r, c)
    t0 = time.perf_counter()
    nbytes = 0
    for _ in range(args.steps):
        nbytes += cpu_reduce_step(L, a, b, m, n, r, c)
    t = time.perf_counter() - t0
    v = nbytes / t / 1e9
    full = log2n == args.log2n
    sample = (f"each step = sum+sqrm+dot+sum-rows+sum-cols on " + ("the whole workload, " if full else f"a bounded sample of 2^{log2n} f32 ({m}x{n}) instead of ")
              + f"2^{args.log2n} f32; oracle/libbaseline.so -O3, single thread (the reference is single threaded, README.md:81; it needs gcc>=14 and "
                f"cannot be compiled in this image, so its loops are restated as ply_ravel compiles them)")
    print(json.dumps({
        "impl": "reference", "metric": "ply map/reduce HBM GB/s", "value": v, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(args),
        "cpu_baseline": {"value": v, "unit": "GB/s", "cores": 1, "kind": "port", "sample": sample, "host_cpu": _cpu_model(), "nproc": os.cpu_count(),
                         "all_cores_sharded_by_hand": cpu_all_cores()},
        "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_name(args):
    return (f"configs[1] reductions on 2^{args.log2n} f32 per GPU: ra::sum, reduce_sqrm, dot + bench-sum-rows/"
            f"bench-sum-cols on {1 << (args.log2n // 2)}x{(1 << args.log2n) >> (args.log2n // 2)}")


def config_of(args):
    """The same object in both arms (the driver compares them)."""
    N = 1 << args.log2n
    m = 1 << (args.log2n // 2)
    return {"workload": workload_name(args), "l2": "inputs (4 GiB per array) larger than L2, no flush needed",
            "bytes_per_step_per_gpu": 4 * N * 6 + 4 * (N // m + m), "sharding": "leading axis, one process per GPU"}


# ---------------------------------------------------------------- clocks

class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) >= 6 and r[2 + i] == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def read_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    except Exception:
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


# ---------------------------------------------------------------- GPU arm

def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    from ra_ra_b200 import abi, cuda, ra

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    lib = cuda.lib
    comm = C.c_void_p()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ident = [None]
        if rank == 0:
            buf = (C.c_ubyte * abi.UNIQUE_ID_BYTES)()
            cuda.check(lib.racu_comm_unique_id(buf))
            ident = [bytes(buf)]
        dist.broadcast_object_list(ident, src=0)
        cuda.check(lib.racu_comm_init_rank(C.byref(comm), ident[0], world, rank))
    ex = cuda.CudaExecutor(local)
    dev = ex.device
    ctx = dict(args=args, torch=torch, dist=dist, ex=ex, lib=lib, cuda=cuda, abi=abi, ra=ra, comm=comm, world=world, rank=rank, local=local)
    if args.workload == "stencil":
        line = run_stencil(ctx, args.n3, args.n0 or args.n3, args.steps, args.warmup)
        if rank == 0:
            print(json.dumps(line))
        if world > 1:
            cuda.check(lib.racu_comm_destroy(comm))
            dist.destroy_process_group()
        return

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        """W untimed + K timed iterations of fn, CUDA events on the current stream, max over ranks (ms per iteration)."""
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item()) / steps
    ctx["timed"] = timed
    peak, peak_source = read_peak()
    ctx["peak"] = peak
    traffic = load_traffic()
    ctx["traffic"] = traffic

    # ---- synthetic inputs: 2^log2n f32 per GPU, uniform [0,1), fixed seeds
    N = 1 << args.log2n
    m = 1 << (args.log2n // 2)
    n = N // m
    g = torch.Generator(device=dev)
    g.manual_seed(0x5EED0011 + rank)
    a_t = torch.rand(N, generator=g, device=dev, dtype=torch.float32)
    g.manual_seed(0x5EED0012 + rank)
    b_t = torch.rand(N, generator=g, device=dev, dtype=torch.float32)
    a, b = ex.wrap(a_t), ex.wrap(b_t)
    A2 = ex.wrap(a_t.view(m, n))
    # sum-rows output and the three scalars (sum, sqrm, dot) share one allocation: one allreduce combines them at N>1
    shared_t = torch.zeros(n + 4, device=dev, dtype=torch.float32)
    rows_t, scal = shared_t[:n], shared_t[n:]
    cols_t = torch.zeros(m, device=dev, dtype=torch.float32)
    rows, cols = ex.wrap(rows_t), ex.wrap(cols_t)

    d_sum, k1 = ra.flatten(ra.as_expr(a), ex=ex)
    d_sqrm, k2 = ra.flatten(ra.sqrm(a), ex=ex)
    d_dot, k3 = ra.flatten(a * b, ex=ex)
    d_rows, k4 = ra.flatten(ra.as_expr(A2), dst=rows(ra.insert(1)), assign=abi.ASSIGN_ADD, ex=ex)
    d_cols, k5 = ra.flatten(ra.as_expr(A2), dst=cols, assign=abi.ASSIGN_ADD, ex=ex)
    sp = cuda.stream_ptr

    dot_events = []  # CUDA events around the dominant kernel (dot: 8 B / element) inside the timed steps

    def step():
        cuda.check(lib.racu_reduce_dev(C.byref(d_sum), abi.RED_SUM, C.c_void_p(scal.data_ptr()), sp()))
        cuda.check(lib.racu_reduce_dev(C.byref(d_sqrm), abi.RED_SUM, C.c_void_p(scal.data_ptr() + 4), sp()))
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        cuda.check(lib.racu_reduce_dev(C.byref(d_dot), abi.RED_SUM, C.c_void_p(scal.data_ptr() + 8), sp()))
        eb.record()
        dot_events.append((ea, eb))
        rows_t.zero_()
        cols_t.zero_()
        cuda.check(lib.racu_map(C.byref(d_rows), sp()))
        cuda.check(lib.racu_map(C.byref(d_cols), sp()))
        if world > 1:  # partials of the sharded leading axis: one allreduce each (sum-cols outputs stay sharded)
            cuda.check(lib.racu_allreduce(comm, C.c_void_p(shared_t.data_ptr()), n + 3, abi.F32, abi.RED_SUM, sp()))

    bytes_step = 4 * N * 6 + 4 * (n + m)  # per GPU, algorithmic
    sampler = ClockSampler(local)
    sampler.start()
    lib.racu_launch_count_reset()
    ms = timed(step, args.steps, args.warmup)
    launches = int(lib.racu_launch_count()) * args.steps // (args.steps + args.warmup)
    sampler.stop_flag = True
    value = bytes_step * world / (ms * 1e-3) / 1e9

    # correctness of what was timed (rank-local, before allreduce scaling): relative error vs fp64
    check = {}
    if world == 1:
        ref = a_t.double()
        check = {"sum_rel_err": abs(float(scal[0]) - float(ref.sum())) / float(ref.sum()),
                 "dot_rel_err": abs(float(scal[2]) - float((ref * b_t.double()).sum())) / float((ref * b_t.double()).sum()),
                 "full-size parity per config": "tests/test_gpu_configs.py (-m gpu)"}
        del ref

    # ---- roofline of the dominant kernel: sflat_fold_inner running dot (8 B / element) + its ply_finish, average launch
    # duration over the timed steps (CUDA events recorded around it on the launching stream, inside the timed region)
    ms_dot = sum(a.elapsed_time(b) for a, b in dot_events[-args.steps:]) / args.steps
    cuda.check(lib.racu_reduce_dev(C.byref(d_dot), abi.RED_SUM, C.c_void_p(scal.data_ptr() + 8), sp()))  # (names the kernel below)
    achieved = 8.0 * N / (ms_dot * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": lib.racu_last_kernel().decode() + "<f32, dot> + ply_finish", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": (traffic.get("dot_f32_2^30") or DOT_TRAFFIC_BYTES) if args.log2n == 30 else None,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel at 2^30 elements (profiles/)",
                "peak_source": peak_source,
                "frac_of_nominal_8TBs": achieved / 8000.0, "bytes_per_launch": 8 * N, "ms_per_launch": ms_dot}

    # ---- e2e: public API (ra.sum / ra.dot / ... on views) from pinned HOST buffers, result read back each step
    e2e = None
    if args.e2e_steps > 0:
        ha = torch.empty(N, dtype=torch.float32).pin_memory()
        hb = torch.empty(N, dtype=torch.float32).pin_memory()
        ha.copy_(a_t)
        hb.copy_(b_t)
        da, db = torch.empty_like(a_t), torch.empty_like(b_t)
        hrows = torch.empty(n, dtype=torch.float32).pin_memory()
        hcols = torch.empty(m, dtype=torch.float32).pin_memory()

        def e2e_step():
            da.copy_(ha, non_blocking=True)
            db.copy_(hb, non_blocking=True)
            va, vb = ex.wrap(da), ex.wrap(db)
            vA2 = ex.wrap(da.view(m, n))
            s0 = ra.sum(va)
            s1 = ra.reduce_sqrm(va)
            s2 = ra.dot(va, vb)
            rows_t.zero_()
            cols_t.zero_()
            rr = rows(ra.insert(1))
            rr += vA2
            cc = cols
            cc += vA2
            hrows.copy_(rows_t, non_blocking=True)
            hcols.copy_(cols_t, non_blocking=True)
            torch.cuda.synchronize()
            return s0, s1, s2
        ms_e2e = timed(e2e_step, args.e2e_steps, 2)
        e2e = {"value": bytes_step * world / (ms_e2e * 1e-3) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": 8 * N,
               "d2h_bytes_per_step": 4 * (n + m) + 12, "ms_per_step": ms_e2e, "steps": args.e2e_steps,
               "bound": "host-to-device copy of both 4 GiB inputs from pinned memory every step (PCIe / NUMA placement of the box); the six device "
                        "passes take %.1f ms of it. At N>1 every rank copies from the same host's memory, so it does not scale with N" % ms}
        del ha, hb, da, db

    # ---- other configs, timed the same way (reported, not the headline)
    extra = {}
    if args.extras and world == 1:
        extra = run_extras(ctx)
    map_sharded = run_map_sharded(ctx) if world > 1 else None

    # ---- the other half of BASELINE.json's metric at this N: laplace3d 1024^3, strong scaling, same process and communicator
    laplace = weak = None
    if args.laplace:
        a_t = b_t = a = b = A2 = None  # (the closures above see the rebinding: the 8 GiB of inputs are released)
        torch.cuda.empty_cache()
        laplace = run_stencil(ctx, args.n3, args.n0 or args.n3, 50, 6, check=True)
        if world == 8:  # the weak-scaling point of BASELINE configs[3]: 2048^3 at 8 GPUs (258 x 2048^2 planes per GPU)
            weak = run_stencil(ctx, 2048, 2048, 20, 4, check=False)

    cpu = cxx = None
    if rank == 0 and world == 1:
        cpu = cpu_baseline(args.cpu_log2n or 28)
        cpu["all_cores_sharded_by_hand"] = cpu_all_cores()
        if args.extras:
            cpu["other_configs"] = cpu_leg_baselines(args.gemm_n)
        cxx = run_cxx_header_bench()
    sampler.join(timeout=2)
    if rank == 0:
        out = {
            "metric": "ply map/reduce HBM GB/s", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": config_of(args),
            "frac_of_measured_hbm_peak": value / world / peak,
            "roofline": roofline, "e2e": e2e, "gpu_launches": launches, "clocks": sampler.summary(), "check": check,
            "extra": extra, "laplace3d": laplace,
        }
        if weak is not None:
            out["laplace3d_weak"] = weak
        if map_sharded is not None:
            out["map_sharded"] = map_sharded
        if cxx is not None:
            out["e2e_cxx"] = cxx
        if cpu is not None:
            out["cpu_baseline"] = cpu
        print(json.dumps(out))
    if world > 1:
        cuda.check(lib.racu_comm_destroy(comm))
        dist.destroy_process_group()


def run_cxx_header_bench():
    """tests/cxx/bench_header.cc (built by __graft_entry__.build()): the C++ host header timed per call and end to end."""
    exe = os.path.join(ROOT, "tests", "cxx", "build", "bench_header_cuda")
    if not os.path.exists(exe):
        return {"unavailable": "tests/cxx/build/bench_header_cuda has not been built (python -c 'import __graft_entry__ as g; g.build()')"}
    try:
        r = subprocess.run([exe, "28", "5"], capture_output=True, text=True, timeout=300)
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)[:200]}


def global_field(torch, lo, hi, plane_shape, dev):
    """Deterministic synthetic field in [-1, 1): planes [lo, hi) of a global grid, the same values whatever the sharding."""
    n1, n2 = plane_shape
    idx = torch.arange(lo * n1 * n2, hi * n1 * n2, device=dev, dtype=torch.int64).view(hi - lo, n1, n2)
    return ((idx * 2654435761 % 2000003).float() / 1000001.5) - 1.0


def bits_checksum(torch, t):
    """Order-independent exact checksum: int64 sum of the 32-bit patterns."""
    return int(t.contiguous().view(torch.int32).sum(dtype=torch.int64))


def run_stencil(ctx, n, n0, steps, warmup, check=False):
    """configs[3]: 7-point sweep (bench/bench-stencil3.cc:55-64) on a GLOBAL f32 n0 x n x n grid, slab-sharded along axis 0 (strong
    scaling: the global grid is fixed as N grows). One step = one sweep + swap. Returns the JSON object (rank 0; None elsewhere)."""
    from ra_ra_b200 import parallel
    args, torch, dist, ex, lib, cuda, abi = (ctx[k] for k in ("args", "torch", "dist", "ex", "lib", "cuda", "abi"))
    comm, world, rank, local = ctx["comm"], ctx["world"], ctx["rank"], ctx["local"]
    dev = ex.device
    sl = parallel.SlabStencil3D((n0, n, n), world, rank)
    push = world > 1 and args.halo == "push"
    gs = sl.global_slice()

    def fresh():
        """(buffers, step function, close): every rank fills its own slab, ghosts included, from the global field."""
        slab = None
        if push:
            slab = cuda.PeerSlab(dist, world, rank, sl.local_shape, abi.F32)  # racu_alloc'ed: exportable with CUDA IPC
            A, An = slab.bufs
            A.copy_(global_field(torch, gs.start, gs.stop, (n, n), dev))
            torch.cuda.synchronize()
            dist.barrier()
            stepper = cuda.PushStepper(sl, slab, abi.STENCIL_3D7, abi.F32)
            state = {"t": 0}

            def step():
                stepper.step()
                state["t"] += 1
            return slab, (lambda: slab.bufs[state["t"] % 2]), step
        A = global_field(torch, gs.start, gs.stop, (n, n), dev)
        An = torch.zeros_like(A)
        bufs = [A, An]
        sweep = cuda.stencil_sweep(abi.STENCIL_3D7, abi.F32)
        exch = parallel.RacuCommExchanger(lib, comm, abi.F32, cuda.stream_ptr, cuda.check) if world > 1 else None
        overlap = cuda.StreamOverlap() if world > 1 and not args.no_overlap else None

        def step():
            sl.step(bufs[0].data_ptr(), bufs[1].data_ptr(), 4, sweep, exch, overlap)
            bufs.reverse()
        return None, (lambda: bufs[0]), step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness first: T steps sharded == T steps on one GPU, bit for bit (order-independent checksum of the bit patterns + one plane)
    chk = None
    if check:
        T = 4
        slab, cur, step = fresh()
        for _ in range(T):
            step()
        torch.cuda.synchronize()
        own = cur()[sl.owned_local()]
        mine = torch.tensor([bits_checksum(torch, own)], device=dev, dtype=torch.int64)
        if world > 1:
            dist.all_reduce(mine, op=dist.ReduceOp.SUM)
        timeouts = slab.timed_out() if slab is not None else None
        if slab is not None:
            slab.close()
        del own, cur, step
        torch.cuda.empty_cache()
        chk = {"steps": T, "checksum": "int64 sum of the f32 bit patterns of every owned plane, summed over the ranks", "sharded": int(mine.item())}
        if rank == 0:  # the same T steps on ONE GPU, whole grid
            A = global_field(torch, 0, n0, (n, n), dev)
            An = torch.zeros_like(A)
            b2 = [A, An]
            sweep = cuda.stencil_sweep(abi.STENCIL_3D7, abi.F32)
            for _ in range(T):
                sweep(b2[0].data_ptr(), b2[1].data_ptr(), (n0, n, n), (n * n, n, 1), 0, 0)
                b2.reverse()
            torch.cuda.synchronize()
            chk["single_gpu"] = bits_checksum(torch, b2[0])
            chk["equal"] = chk["single_gpu"] == chk["sharded"]
            chk["halo_timeouts"] = timeouts
            del A, An, b2
            torch.cuda.empty_cache()
        barrier()

    slab, cur, step = fresh()
    for _ in range(warmup + (warmup % 2)):
        step()
    barrier()
    nsteps = steps + (steps % 2)
    lib.racu_launch_count_reset()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(nsteps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    barrier()
    sampler.stop_flag = True
    ms = float(ms.item()) / nsteps
    launches = int(lib.racu_launch_count())
    cells = float(n0) * float(n) ** 2
    peak, _ = read_peak()
    gcells = cells / (ms * 1e-3) / 1e9
    sampler.join(timeout=2)
    line = None
    if rank == 0:
        tr = ctx.get("traffic", {}).get(f"stencil3d7_f32_{n}^3") if world == 1 and n0 == n else None
        line = ({
            "metric": "laplace3d Gcells/s", "value": gcells, "unit": "Gcells/s", "n_gpus": world, "steps": nsteps,
            "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if n == 1024 or world == 1 else "weak (2048^3 at 8 GPUs: the slab per GPU is 258 x 2048^2)",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[3] 7-point stencil f32 {n0}x{n}x{n} global grid, slab-sharded along axis 0, " + (
                                   "one kernel per step: sweep + halo push into the neighbours' ghost planes over NVLink peer memory "
                                   "(CUDA IPC), flag lock step" if push else
                                   ("one GPU, no exchange" if world == 1 else f"NCCL halo exchange per step, overlap={'off' if args.no_overlap else 'interior/edges on two streams'}")) +
                                   ", launched step by step from the host",
                       "l2": "%.1f GiB per buffer and GPU, larger than L2" % (4 * cells / world / GIB)},
            "halo_timeouts": slab.timed_out() if push else None,
            "roofline": {"bound": "hbm", "kernel": lib.racu_last_kernel().decode(), "achieved": 8 * cells / world / (ms * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": 8 * cells / world / (ms * 1e-3) / 1e9 / peak, "traffic": tr,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one sweep, ncu --set full (profiles/)" if tr else None,
                         "note": "per GPU: 8 B/cell algorithmic x cells per GPU / step time (includes exchange when N>1)"},
            "gpu_launches": launches, "clocks": sampler.summary()})
        if chk is not None:
            line["check"] = chk
    if slab is not None:
        slab.close()
    del cur, step
    torch.cuda.empty_cache()
    if world == 1 and n0 == n and line is not None:
        line["two_steps_per_pass"] = run_stencil_two_per_pass(ctx, n, nsteps, ms)
    return line


def run_stencil_two_per_pass(ctx, n, nsteps, ms_single):
    """The same T steps through racu_stencil_steps: the benchmark's loop (bench/bench-stencil3.cc:55-64 + swap, :117-121) handed to the library
    as ONE call, which takes the steps two per pass (state t+1 only in shared memory: 8 B/cell of HBM traffic for two steps). One GPU.
    check: 10 steps leave BOTH buffers bit-identical to 10 single steps (int64 sums of the bit patterns)."""
    torch, ex, lib, cuda = (ctx[k] for k in ("torch", "ex", "lib", "cuda"))
    dev = ex.device
    sweep = cuda.stencil_sweep(ctx["abi"].STENCIL_3D7, ctx["abi"].F32)
    sums = []
    scratch = torch.empty((n, n, n), device=dev, dtype=torch.float32)
    for fused in (False, True):
        A = global_field(torch, 0, n, (n, n), dev)
        An = torch.zeros_like(A)
        if fused:
            cuda.stencil_steps(A, An, 10, scratch=scratch)
        else:
            b2 = [A, An]
            for _ in range(10):
                sweep(b2[0].data_ptr(), b2[1].data_ptr(), (n, n, n), (n * n, n, 1), 0, 0)
                b2.reverse()
        torch.cuda.synchronize()
        sums.append((bits_checksum(torch, A), bits_checksum(torch, An)))
    cuda.stencil_steps(A, An, 6, scratch=scratch)   # warm-up
    torch.cuda.synchronize()
    lib.racu_launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    cuda.stencil_steps(A, An, nsteps, scratch=scratch)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / nsteps
    launches = int(lib.racu_launch_count())
    peak, _ = read_peak()
    cells = float(n) ** 3
    tr = ctx.get("traffic", {}).get(f"stencil3d7_two_steps_f32_{n}^3")
    out = {"metric": "laplace3d Gcells/s", "value": cells / (ms * 1e-3) / 1e9, "unit": "Gcells/s per step", "steps": nsteps, "ms_per_step": ms,
           "speedup_over_single_steps": ms_single / ms, "gpu_launches": launches,
           "kernel": "stencil_tma<3D7, two steps> (%d passes) + copy_faces + 2 single steps" % (nsteps // 2 - 1),
           "roofline": {"bound": "hbm", "achieved": 8 * cells / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": 8 * cells / (ms * 1e-3) / 1e9 / peak,
                        "traffic": tr,
                        "note": "achieved = the 8 B/cell/step a single-step sweep needs / time per step: above 1.0 because a pass reads and writes the grid once for TWO steps "
                                "(the pass itself moves 8 B/cell: its own HBM fraction is half of this)"},
           "check": {"steps": 10, "single_steps": list(sums[0]), "two_per_pass": list(sums[1]), "equal_both_buffers": sums[0] == sums[1]}}
    del A, An, scratch
    torch.cuda.empty_cache()
    return out


def run_map_sharded(ctx):
    """configs[0] at N>1: B += A*C + v on f32 8192^2 PER GPU, rows sharded (weak scaling), no collective: every rank runs its own slab
    through racu_map; checked bit for bit against the same ops evaluated op by op with torch."""
    args, torch, dist, ex, lib, cuda, abi, ra, timed = (ctx[k] for k in ("args", "torch", "dist", "ex", "lib", "cuda", "abi", "ra", "timed"))
    world, rank, dev, peak = ctx["world"], ctx["rank"], ctx["ex"].device, ctx["peak"]
    n = 8192
    g = torch.Generator(device=dev)
    ts = []
    for seed in (1, 2, 3):
        g.manual_seed(0x5EED0000 + seed + 16 * rank)
        ts.append(torch.rand(n, n, generator=g, device=dev, dtype=torch.float32) * 2 - 1)
    A, B, Cc = (ex.wrap(t) for t in ts)
    g.manual_seed(0x5EED0004 + 16 * rank)
    v_t = torch.rand(n, generator=g, device=dev, dtype=torch.float64) * 2 - 1
    v = ex.wrap(v_t)
    B0 = ts[1].clone()
    d, keep = ra.flatten(A * Cc + v, dst=B, assign=abi.ASSIGN_ADD, ex=ex)
    cuda.check(lib.racu_map(C.byref(d), cuda.stream_ptr()))
    ok = torch.tensor([int(torch.equal(ts[1], (B0.double() + ((ts[0] * ts[2]).double() + v_t[:, None])).float()))], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    ms = timed(lambda: cuda.check(lib.racu_map(C.byref(d), cuda.stream_ptr())), 20, 3)
    gbs = 16.0 * n * n * world / (ms * 1e-3) / 1e9
    out = {"metric": "ply map HBM GB/s", "value": gbs, "unit": "GB/s", "n_gpus": world, "scaling": "weak", "ms_per_step": ms,
           "config": {"workload": "configs[0] B += A*C + v(double, prefix broadcast) f32 8192^2 per GPU, rows sharded, no collective"},
           "frac_of_measured_hbm_peak_per_gpu": gbs / world / peak, "kernel": lib.racu_last_kernel().decode(),
           "check": {"bit exact vs the same ops in torch, every rank": bool(ok.item())}}
    del ts, A, B, Cc, B0
    torch.cuda.empty_cache()
    return out


def run_extras(ctx):
    """All five configs on one GPU, same timing rules; each entry names its kernel and its fraction of the measured HBM peak."""
    args, torch, ex, lib, cuda, ra, abi, timed, peak = (ctx[k] for k in ("args", "torch", "ex", "lib", "cuda", "ra", "abi", "timed", "peak"))
    traffic = ctx.get("traffic", {})
    dev = ex.device
    out = {}
    g = torch.Generator(device=dev)
    # configs[0]: B += A*C (+ v row broadcast) on f32 8192^2, 16 B/element
    n = 8192
    ts = []
    for seed in (1, 2, 3):
        g.manual_seed(0x5EED0000 + seed)
        ts.append(torch.rand(n, n, generator=g, device=dev, dtype=torch.float32) * 2 - 1)
    A, B, Cc = (ex.wrap(t) for t in ts)
    g.manual_seed(0x5EED0004)
    v = ex.wrap(torch.rand(n, generator=g, device=dev, dtype=torch.float64) * 2 - 1)
    d1, k1 = ra.flatten(A * Cc, dst=B, assign=abi.ASSIGN_ADD, ex=ex)
    d2, k2 = ra.flatten(A * Cc + v, dst=B, assign=abi.ASSIGN_ADD, ex=ex)
    for name, d in (("map B+=A*C f32 8192^2", d1), ("map B+=A*C+v(double, prefix broadcast) f32 8192^2", d2)):
        ms = timed(lambda: cuda.check(lib.racu_map(C.byref(d), cuda.stream_ptr())), 20, 3)
        gbs = 16.0 * n * n / (ms * 1e-3) / 1e9
        out[name] = {"ms": ms, "GB/s": gbs, "frac_of_measured_hbm_peak": gbs / peak, "kernel": lib.racu_last_kernel().decode(),
                     "l2": "working set 768 MiB > L2", "traffic": traffic.get("map_f32_8192^2")}
    del ts, A, B, Cc
    # configs[3]: 7-point stencil f32 1024^3, 8 B/cell/step
    for n3 in (1024,):
        A = torch.empty(n3, n3, n3, device=dev, dtype=torch.float32)
        g.manual_seed(0x5EED0041)
        A.uniform_(-1, 1, generator=g)
        An = torch.zeros_like(A)
        bufs = [A, An]
        s = abi.StencilDesc()
        s.kind, s.dtype, s.rank = abi.STENCIL_3D7, abi.F32, 3
        for k in range(3):
            s.len[k] = n3
            s.src_stride[k] = s.dst_stride[k] = A.stride(k)

        def sweep():
            s.src, s.dst = bufs[0].data_ptr(), bufs[1].data_ptr()
            cuda.check(lib.racu_stencil(C.byref(s), cuda.stream_ptr()))
            bufs.reverse()  # std::swap(A.cp, Anext.cp)
        ms = timed(sweep, 10, 4)
        cells = float(n3) ** 3
        out[f"stencil3d7 f32 {n3}^3"] = {"ms_per_step": ms, "Gcells/s": cells / (ms * 1e-3) / 1e9,
                                         "GB/s_algorithmic": 8 * cells / (ms * 1e-3) / 1e9,
                                         "frac_of_measured_hbm_peak": 8 * cells / (ms * 1e-3) / 1e9 / peak,
                                         "kernel": lib.racu_last_kernel().decode(), "traffic": traffic.get(f"stencil3d7_f32_{n3}^3")}
        del A, An, bufs
    torch.cuda.empty_cache()
    # configs[2]: strided views on f64 1024^3 (transpose, reverse, iota-slice assignment), 16 B per touched element; pure data movement
    n3 = 1024
    A_t = torch.empty(n3, n3, n3, device=dev, dtype=torch.float64)
    A = ex.wrap(A_t)
    A.assign(ra._0 * (n3 * n3) + ra._1 * n3 + ra._2)  # A[i,j,k] = i n^2 + j n + k: a misplaced element is detectable
    B_t = torch.zeros_like(A_t)
    B = ex.wrap(B_t)
    views = [("B = transpose(A,{2,1,0})", ra.transpose(A, (2, 1, 0)), B, n3 ** 3, lambda: torch.equal(B_t, A_t.permute(2, 1, 0))),
             ("B = transpose(A,{0,2,1})", ra.transpose(A, (0, 2, 1)), B, n3 ** 3, lambda: torch.equal(B_t, A_t.permute(0, 2, 1))),
             ("B = reverse(A,2)", ra.reverse(A, 2), B, n3 ** 3, lambda: torch.equal(B_t, A_t.flip(2))),
             ("B = reverse(A,0)", ra.reverse(A, 0), B, n3 ** 3, lambda: torch.equal(B_t, A_t.flip(0))),
             ("B(iota(n/2,0,2), all, iota(n,n-1,-1)) = A(iota(n/2), all, all)", A(ra.iota(n3 // 2), ra.all, ra.all),
              B(ra.iota(n3 // 2, 0, 2), ra.all, ra.iota(n3, n3 - 1, -1)), n3 ** 3 // 2, lambda: torch.equal(B_t[0::2].flip(2), A_t[: n3 // 2]))]
    checks = {}
    for name, src, dst, elems, ok in views:
        d, keep = ra.flatten(ra.as_expr(src), dst=dst, assign=abi.ASSIGN_SET, ex=ex)
        ms = timed(lambda: cuda.check(lib.racu_map(C.byref(d), cuda.stream_ptr())), 5, 3)
        gbs = 16.0 * elems / (ms * 1e-3) / 1e9
        out["views f64 1024^3: " + name] = {"ms": ms, "GB/s": gbs, "frac_of_measured_hbm_peak": gbs / peak, "kernel": lib.racu_last_kernel().decode()}
        checks[name] = bool(ok())  # element for element against the same view taken with torch indexing
    out["views f64 1024^3: check"] = {"every assignment equals the index arithmetic (element for element)": checks}
    del A_t, B_t, A, B, views
    torch.cuda.empty_cache()
    # configs[4]: dense contraction and gemv / gevm (bench-gemm / bench-gemv) at BASELINE's 16384^2: f64 on FP64 tensor-core MMA, f32 as 3xTF32 on tcgen05
    ng = args.gemm_n
    for dt, tname, es in ((torch.float64, "f64", 8), (torch.float32, "f32", 4)):
        g.manual_seed(0x5EED0051)
        a_t = torch.rand(ng, ng, generator=g, device=dev, dtype=dt) * 2 - 1
        g.manual_seed(0x5EED0052)
        b_t = torch.rand(ng, ng, generator=g, device=dev, dtype=dt) * 2 - 1
        c_t = torch.zeros(ng, ng, device=dev, dtype=dt)
        gd = abi.GemmDesc()
        gd.dtype, gd.M, gd.N, gd.K = (abi.F64 if es == 8 else abi.F32), ng, ng, ng
        gd.a, gd.lda, gd.b, gd.ldb, gd.c, gd.ldc = a_t.data_ptr(), ng, b_t.data_ptr(), ng, c_t.data_ptr(), ng
        ms = timed(lambda: cuda.check(lib.racu_gemm(C.byref(gd), cuda.stream_ptr())), 3, 1)
        kern = lib.racu_last_kernel().decode()
        c_t.zero_()
        cuda.check(lib.racu_gemm(C.byref(gd), cuda.stream_ptr()))
        ref = a_t[:64].double() @ b_t.double()
        err = float((c_t[:64].double() - ref).abs().max() / (a_t[:64].double().abs() @ b_t.double().abs()).max())
        # the library GEMM on the same box as the practical ceiling (cuBLAS through torch; fp32 without TF32)
        torch.backends.cuda.matmul.allow_tf32 = False
        ms_lib = timed(lambda: torch.matmul(a_t, b_t, out=c_t), 2, 1)
        out[f"gemm {tname} {ng}^3"] = {"ms": ms, "TFLOP/s": 2.0 * ng ** 3 / (ms * 1e-3) / 1e12, "kernel": kern,
                                       "max_err_over_sum_abs_products": err, "bound": "tensor pipe (f64: FP64 MMA; f32: 3xTF32 on tcgen05, TMEM accumulator)",
                                       "cublas_same_box_TFLOP/s": 2.0 * ng ** 3 / (ms_lib * 1e-3) / 1e12}
        x_t, y_t = torch.rand(ng, device=dev, dtype=dt), torch.zeros(ng, device=dev, dtype=dt)
        for trans, nm in ((0, "gemv"), (1, "gevm")):
            ms = timed(lambda: cuda.check(lib.racu_gemv(gd.dtype, trans, ng, ng, a_t.data_ptr(), ng, x_t.data_ptr(), 1, y_t.data_ptr(), 1, cuda.stream_ptr())), 20, 3)
            gbs = es * ng * ng / (ms * 1e-3) / 1e9
            out[f"{nm} {tname} {ng}^2"] = {"ms": ms, "GB/s": gbs, "frac_of_measured_hbm_peak": gbs / peak, "kernel": lib.racu_last_kernel().decode(),
                                           "l2": "matrix %d MiB > L2" % (es * ng * ng >> 20), "traffic": traffic.get(f"{nm}_{tname}_{ng}^2")}
        del a_t, b_t, c_t, x_t, y_t
    return out


if __name__ == "__main__":
    main()
