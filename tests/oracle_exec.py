"""Executor that runs racu descriptors on HOST memory through the CPU oracle (oracle/liboracle.so).
Test infrastructure only: the checker the CUDA path is compared against."""
import ctypes as C
import os

import numpy as np

from ra_ra_b200 import abi, ra

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        L.oracle_last_error_string.restype = C.c_char_p
        L.oracle_agree.argtypes = [C.POINTER(abi.Desc), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
        L.oracle_map.argtypes = [C.POINTER(abi.Desc)]
        L.oracle_reduce.argtypes = [C.POINTER(abi.Desc), C.c_int, C.c_void_p]
        L.oracle_stencil_raw.argtypes = [C.POINTER(abi.StencilDesc)]
        L.oracle_gemm_raw.argtypes = [C.POINTER(abi.GemmDesc)]
        _lib = L
    return _lib


_prep = None


def _prepare(d, redop=0):
    """RACU_ORACLE_PREPARE=1 (set by tests/test_oracle_golden.py): every descriptor the oracle executes is also handed to the product's
    racu_prepare - plan + run-time specialisation, no GPU, nothing launched (include/racu.h) - so that the CPU run of the known-answer cases
    leaves the kernels their GPU run needs in the specialiser's disk cache (ra_ra_b200/jit_cache travels with the tree)."""
    global _prep
    if not os.environ.get("RACU_ORACLE_PREPARE"):
        return
    if _prep is None:
        try:
            from ra_ra_b200 import build
            _prep = abi.bind(C.CDLL(build.build_lib()))
        except Exception:  # noqa: BLE001
            _prep = False
    if _prep:
        k = C.c_char_p()
        _prep.racu_prepare(C.byref(d), int(redop), C.byref(k))


class OracleDim(C.Structure):
    _fields_ = [("len", C.c_int64), ("step", C.c_int64)]


class OracleSub(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n", C.c_int32), ("len", C.c_int64), ("org", C.c_int64), ("step", C.c_int64)]


def _check(st):
    if st != abi.OK:
        raise ra.RaError(st, lib().oracle_last_error_string().decode())


class OracleExecutor:
    name = "oracle"

    def empty(self, shape, dtype):
        arr = np.empty(shape, dtype=ra.RACU2NP[dtype])
        return self.wrap(arr)

    def wrap(self, arr):
        """View over an existing C-contiguous numpy array (shares memory)."""
        assert arr.flags.c_contiguous
        steps = [s // arr.itemsize for s in arr.strides]
        return ra.View(self, arr, arr.ctypes.data, ra.NP2RACU[arr.dtype], list(zip(arr.shape, steps)))

    def from_numpy(self, arr):
        return self.wrap(np.array(arr, order="C", copy=True))

    def to_numpy(self, v):
        out = np.empty(v.shape, dtype=ra.RACU2NP[v.dtype])
        self.wrap(out).assign(v)
        return out

    def map(self, d):
        _prepare(d)
        _check(lib().oracle_map(C.byref(d)))

    def agree(self, d):
        rank = C.c_int32()
        lens = (C.c_int64 * abi.MAX_RANK)()
        _check(lib().oracle_agree(C.byref(d), C.byref(rank), lens))
        return [lens[k] for k in range(rank.value)]

    def reduce(self, d, redop, rtype):
        out = np.zeros(1, dtype=ra.RACU2NP[rtype])
        _prepare(d, redop)
        _check(lib().oracle_reduce(C.byref(d), redop, out.ctypes.data))
        return out[0]


def _oracle_reduce_into(self, d, redop, ptr):
    _prepare(d, redop)
    _check(lib().oracle_reduce(C.byref(d), redop, C.c_void_p(ptr)))


OracleExecutor.reduce_into = _oracle_reduce_into


def oracle_reverse(dims, k):
    dv = (OracleDim * len(dims))(*[OracleDim(l, s) for l, s in dims])
    off = C.c_int64()
    _check(lib().oracle_reverse(dv, len(dims), k, C.byref(off)))
    return off.value, [[d.len, d.step] for d in dv]


def oracle_transpose(dims, s):
    sv = (OracleDim * len(dims))(*[OracleDim(l, st) for l, st in dims])
    dv = (OracleDim * abi.MAX_RANK)()
    sa = (C.c_int32 * len(s))(*s)
    r = C.c_int32()
    _check(lib().oracle_transpose(sv, len(dims), sa, dv, C.byref(r)))
    return [[dv[k].len, dv[k].step] for k in range(r.value)]


def oracle_from(dims, subs):
    """subs: list of ('scalar', i) | ('iota', n, o, s) | ('dots', n) | ('insert', n)."""
    sv = (OracleDim * max(1, len(dims)))(*[OracleDim(l, st) for l, st in dims])
    ss = (OracleSub * max(1, len(subs)))()
    for q, s in enumerate(subs):
        if s[0] == "scalar":
            ss[q] = OracleSub(0, 0, 0, s[1], 0)
        elif s[0] == "iota":
            ss[q] = OracleSub(1, 0, s[1], s[2], s[3])
        elif s[0] == "dots":
            ss[q] = OracleSub(2, s[1], 0, 0, 0)
        else:
            ss[q] = OracleSub(3, s[1], 0, 0, 0)
    bv = (OracleDim * abi.MAX_RANK)()
    r = C.c_int32()
    off = C.c_int64()
    _check(lib().oracle_from(sv, len(dims), ss, len(subs), bv, C.byref(r), C.byref(off)))
    return off.value, [[bv[k].len, bv[k].step] for k in range(r.value)]
