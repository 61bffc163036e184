"""Executor over the planner's CPU simulator (ra_ra_b200/csrc/sim/libracu_sim.so): make_plan() + the kernels' own
per-vector bodies, walked sequentially on HOST memory. Test infrastructure only; checks the planner without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle_exec import OracleExecutor
from ra_ra_b200 import abi, ra

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "ra_ra_b200", "csrc")
_lib = None


def build():
    out = os.path.join(CSRC, "sim", "libracu_sim.so")
    srcs = [os.path.join(CSRC, "sim", "plan_sim.cc"), os.path.join(CSRC, "plan.cc")]
    deps = srcs + [os.path.join(CSRC, n) for n in ("kcore.h", "kbody.h", "plan.h")] + [os.path.join(ROOT, "include", "racu.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++20", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas",
                        "-o", out] + srcs, check=True)
    return out


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.racu_sim_last_error.restype = C.c_char_p
        L.racu_sim_run.argtypes = [C.POINTER(abi.Desc), C.c_int, C.c_void_p]
        L.racu_sim_plan_info.argtypes = [C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
        _lib = L
    return _lib


def plan_info():
    a = (C.c_int32 * 16)()
    b = (C.c_int64 * 8)()
    lib().racu_sim_plan_info(a, b)
    keys = ["mode", "wide", "rankA", "rankB", "nstage", "depth", "U", "nchunks", "grid_x", "grid_y", "smem", "need_finish",
            "nins", "empty", "nvec", "nopnd"]
    d = dict(zip(keys, a))
    d.update(dict(zip(["len0", "nvecA", "nB", "nk", "chunk_len"], b)))
    return d


class SimExecutor(OracleExecutor):
    name = "sim"

    def map(self, d):
        st = lib().racu_sim_run(C.byref(d), 0, None)
        if st != abi.OK:
            raise ra.RaError(st, lib().racu_sim_last_error().decode())

    def reduce(self, d, redop, rtype):
        out = np.zeros(2, dtype=ra.RACU2NP[rtype])
        st = lib().racu_sim_run(C.byref(d), redop, out.ctypes.data)
        if st != abi.OK:
            raise ra.RaError(st, lib().racu_sim_last_error().decode())
        return out[0]

    def reduce_into(self, d, redop, ptr):
        st = lib().racu_sim_run(C.byref(d), redop, C.c_void_p(ptr))
        if st != abi.OK:
            raise ra.RaError(st, lib().racu_sim_last_error().decode())
