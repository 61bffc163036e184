"""Run-time specialiser without a GPU: racu_prepare plans a descriptor and has NVRTC compile the specialised kernel for
sm_100a from the sources embedded in libracu.so (csrc/jit.cc). Host pointers are fine - nothing is launched."""
import ctypes as C
import os

import pytest

from jit_cases import CASES, make_inputs
from oracle_exec import OracleExecutor
from ra_ra_b200 import abi, build, ra


@pytest.fixture(scope="module")
def lib():
    return abi.bind(C.CDLL(build.build_lib()))


def _prepare(lib, name):
    kind, make = CASES[name]
    ex = OracleExecutor()
    e, dst, assign, redop = make(make_inputs(ex))
    d, keep = ra.flatten(e, dst=dst, assign=assign, ex=ex) if dst is not None else ra.flatten(e, ex=ex)
    k = C.c_char_p()
    st = lib.racu_prepare(C.byref(d), redop, C.byref(k))
    assert st == abi.OK, lib.racu_last_error_string()
    return k.value.decode()


@pytest.mark.parametrize("name", sorted(CASES))
def test_prepare_specialises(name, lib):
    """Also warms the default disk cache (ra_ra_b200/jit_cache), which travels with the tree to the GPU box."""
    kernel = _prepare(lib, name)
    want = {"map": "jit_map", "fold": "jit_fold_inner"}.get(CASES[name][0])
    assert kernel == want or (want is None and kernel in ("jit_fold_inner", "jit_fold_outer")), kernel


def test_disk_cache_and_table_programs(lib, tmp_path, monkeypatch):
    monkeypatch.setenv("RACU_JIT_CACHE", str(tmp_path))
    assert _prepare(lib, "int_ops") == "jit_map"
    cubins = [f for f in os.listdir(tmp_path) if f.endswith(".cubin")]
    assert len(cubins) == 1 and os.path.getsize(tmp_path / cubins[0]) > 1000
    # a program of the pre-compiled table is not re-specialised
    ex = OracleExecutor()
    q = make_inputs(ex)
    d, keep = ra.flatten(q["A"] * q["C"], dst=q["B"], assign=abi.ASSIGN_ADD, ex=ex)
    k = C.c_char_p()
    assert lib.racu_prepare(C.byref(d), 0, C.byref(k)) == abi.OK and k.value == b"sflat_map"
    assert len(os.listdir(tmp_path)) == 1
    # RACU_NO_JIT keeps the interpreted kernels
    monkeypatch.setenv("RACU_NO_JIT", "1")
    assert _prepare(lib, "int_ops").startswith("ply_map")
