"""The C-ABI library loads without a GPU and exports every symbol include/racu.h declares."""
import ctypes as C
import os
import re

import numpy as np

from ra_ra_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from ra_ra_b200 import build
    return C.CDLL(build.build_lib())


def test_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "racu.h")).read()
    declared = set(re.findall(r"\b(racu_[a-z0-9_]+)\s*\(", hdr))
    assert declared == {name for name, _, _ in abi.SYMBOLS}, declared ^ {name for name, _, _ in abi.SYMBOLS}
    lib = _lib()
    for name in declared:
        assert hasattr(lib, name), name
    abi.bind(lib)
    assert lib.racu_version() == abi.VERSION


def test_struct_layout_matches_header():
    # sizes implied by include/racu.h on LP64
    assert C.sizeof(abi.Operand) == 16 + 8 + 8 * 8 * 2
    assert C.sizeof(abi.Ins) == 4
    assert C.sizeof(abi.Desc) == 16 + abi.MAX_OPND * C.sizeof(abi.Operand) + abi.MAX_INS * 4
    assert C.sizeof(abi.StencilDesc) == 16 + 16 + 9 * 8 + 16
    assert C.sizeof(abi.GemmDesc) == 8 + 3 * 8 + 6 * 8


def test_agree_needs_no_device():
    """Shape agreement (Match, ra/expr.hh:532-608) is host logic inside the library."""
    lib = abi.bind(_lib())
    d = abi.Desc()
    d.nopnd, d.nins, d.dst = 2, 1, 0
    buf = np.zeros(24)
    for o, shape in enumerate([(3, 2, 4), (3, 2)]):
        p = d.opnd[o]
        p.kind, p.dtype, p.rank = abi.MEM, abi.F64, len(shape)
        p.base.ptr = buf.ctypes.data
        for k, n in enumerate(shape):
            p.len[k] = n
    rank = C.c_int32()
    lens = (C.c_int64 * abi.MAX_RANK)()
    assert lib.racu_agree(C.byref(d), C.byref(rank), lens) == abi.OK
    assert rank.value == 3 and list(lens[:3]) == [3, 2, 4]
    d.opnd[1].len[1] = 3
    assert lib.racu_agree(C.byref(d), C.byref(rank), lens) == abi.ERR_SHAPE
    assert b"Bad shapes" in lib.racu_last_error_string()


def test_malformed_descriptors_are_rejected_before_use():
    """Every entry point validates counts and indices before anything dereferences opnd[dst] / opnd[ins.arg] (no GPU needed:
    validation comes first)."""
    lib = abi.bind(_lib())
    buf = np.zeros(16, np.float32)

    def good():
        d = abi.Desc()
        d.nopnd, d.nins, d.dst, d.assign = 2, 1, 0, abi.ASSIGN_SET
        for o in range(2):
            p = d.opnd[o]
            p.kind, p.dtype, p.rank = abi.MEM, abi.F32, 1
            p.base.ptr = buf.ctypes.data
            p.len[0], p.stride[0] = 4, 1
        d.ins[0].op, d.ins[0].type, d.ins[0].arg = abi.OP_PUSH, abi.F32, 1
        return d
    cases = []
    d = good(); d.dst = 7; cases.append(d)
    d = good(); d.dst = abi.MAX_OPND + 5; cases.append(d)
    d = good(); d.nopnd = abi.MAX_OPND + 1; cases.append(d)
    d = good(); d.nopnd = 0; cases.append(d)
    d = good(); d.nins = abi.MAX_INS + 1; cases.append(d)
    d = good(); d.nins = 0; cases.append(d)
    d = good(); d.ins[0].arg = 255; cases.append(d)
    d = good(); d.ins[0].type = 9; cases.append(d)
    d = good(); d.opnd[1].rank = abi.MAX_RANK + 1; cases.append(d)
    d = good(); d.opnd[1].kind = 17; cases.append(d)
    # stencil-looking garbage: >= 8 instructions with a destination index out of range
    d = good(); d.nins = 12; d.dst = 200; cases.append(d)
    for d in cases:
        assert lib.racu_map(C.byref(d), None) == abi.ERR_BAD_DESC
        assert lib.racu_prepare(C.byref(d), 0, None) == abi.ERR_BAD_DESC
    out = np.zeros(2)
    d = good()  # reductions need dst = -1
    assert lib.racu_reduce(C.byref(d), abi.RED_SUM, out.ctypes.data, None) == abi.ERR_BAD_DESC
    assert lib.racu_reduce_dev(C.byref(d), abi.RED_SUM, out.ctypes.data, None) == abi.ERR_BAD_DESC
