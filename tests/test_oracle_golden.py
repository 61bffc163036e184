"""Pins the CPU oracle (oracle/ply_oracle.cc) against the reference's own known-answer tests (tests/kat.py lists
them with file:line), and checks the oracle's view algebra against the Python host mirror."""
import os

import numpy as np
import pytest

import kat
from oracle_exec import OracleExecutor, oracle_from, oracle_reverse, oracle_transpose
from ra_ra_b200 import abi, ra


@pytest.mark.parametrize("case", kat.CASES, ids=lambda f: f.__name__)
def test_kat_on_oracle(case, monkeypatch):
    """(With RACU_WARM_JIT=1 every descriptor is also handed to the product's racu_prepare, which leaves the specialised kernels of these
    cases in the disk cache for their GPU run, tests/test_gpu_kat.py: oracle_exec._prepare, tests/test_warm_jit.py.)"""
    if os.environ.get("RACU_WARM_JIT"):
        monkeypatch.setenv("RACU_ORACLE_PREPARE", "1")
    case(OracleExecutor())


@pytest.fixture
def fixed_traversal():
    from oracle_exec import lib
    lib().oracle_set_traversal(1)
    yield
    lib().oracle_set_traversal(0)


@pytest.mark.parametrize("case", kat.CASES, ids=lambda f: f.__name__)
def test_kat_on_oracle_ply_fixed(case, fixed_traversal):
    """The reference runs its traversal tests through ply_ravel AND ply_fixed (TEST(plier), test/ply.cc:35-44); so does the oracle:
    the same known answers through the restated ply_fixed (ra/ply.hh:90-153: no collapsing of axes)."""
    case(OracleExecutor())


@pytest.mark.parametrize("seed", range(30))
def test_ply_ravel_and_ply_fixed_agree_on_random_expressions(seed):
    from oracle_exec import lib
    from randexpr import DT, SHAPES, random_expr, random_view
    outs = []
    for fixed in (0, 1):
        lib().oracle_set_traversal(fixed)
        try:
            rng = np.random.default_rng(2000 + seed)
            ex = OracleExecutor()
            shape = SHAPES[seed % len(SHAPES)]
            dst, _ = random_view(rng, ex, shape, DT[rng.integers(0, 4)])
            e = random_expr(rng, ex, shape, 2)
            try:
                dst += e
            except ra.RaError as err:
                assert err.status == abi.ERR_UNSUPPORTED, err
            outs.append(np.array(dst.owner, copy=True))
        finally:
            lib().oracle_set_traversal(0)
    assert np.array_equal(outs[0], outs[1], equal_nan=True)


def test_view_algebra_matches_oracle():
    ex = OracleExecutor()
    a = ex.wrap(np.zeros((3, 4, 5, 6)))
    for k in range(4):
        off, dims = oracle_reverse(a.dims, k)
        b = ra.reverse(a, k)
        assert (b.ptr - a.ptr) // 8 == off and b.dims == dims
    for s in ((1, 0, 2, 3), (3, 2, 1, 0), (0, 0, 1, 1), (2, 0, 3, 1), (0, 1, 1, 4)):
        assert ra.transpose(a, s).dims == oracle_transpose(a.dims, s)
    cases = [
        ((2, ra.all, ra.iota(3, 1), 5), [("scalar", 2), ("dots", 1), ("iota", 3, 1, 1), ("scalar", 5)]),
        ((ra.iota(2, 2, -1), ra.dots(), 0), [("iota", 2, 2, -1), ("dots", -1), ("scalar", 0)]),
        ((ra.insert(2), 1), [("insert", 2), ("scalar", 1)]),
        ((ra.dots(2), ra.insert(1), ra.iota(0, 0)), [("dots", 2), ("insert", 1), ("iota", 0, 0, 1)]),
        ((ra.all, ra.iota(4, 0, 0)), [("dots", 1), ("iota", 4, 0, 0)]),
    ]
    for subs, osubs in cases:
        b = a(*subs)
        off, dims = oracle_from(a.dims, osubs)
        assert (b.ptr - a.ptr) // 8 == off and b.dims == dims, (subs, b.dims, dims)
    with pytest.raises(ra.RaError):
        oracle_from(a.dims, [("scalar", 3)])
    with pytest.raises(ra.RaError):
        oracle_from(a.dims, [("iota", 3, 2, 1)])


def test_oracle_sequential_float_sum_is_sequential():
    """ra/ra.hh:351-352: the accumulator has the element type and the fold is in row-major order."""
    ex = OracleExecutor()
    x = np.random.default_rng(3).uniform(0, 1, 100003).astype(np.float32)
    c = np.float32(0)
    for v in x:
        c = np.float32(c + v)
    assert ra.sum(ex.wrap(x)) == c
    c = np.float32(0)
    for v in x:
        c = np.float32(c + np.float32(v * v))
    assert ra.reduce_sqrm(ex.wrap(x)) == c


def test_oracle_mixed_type_assign_narrowing():
    """float B += double expr computes in double then narrows."""
    ex = OracleExecutor()
    B = ex.from_numpy(np.array([0.1, 0.2], dtype=np.float32))
    v = np.array([1e-9, 1.0 / 3.0])
    b0 = B.numpy()
    B += np.float32(3.0) * B + v
    expect = (b0.astype(np.float64) + ((np.float32(3.0) * b0).astype(np.float64) + v)).astype(np.float32)
    assert np.array_equal(B.numpy(), expect)


def test_program_validation():
    ex = OracleExecutor()
    a = ex.from_numpy(np.zeros(3, dtype=np.float32))
    d, keep = ra.flatten(ra.as_expr(a) + 1.0, dst=a)
    d.ins[d.nins - 1].type = abi.F32  # ADD typed f32 over f64 operands
    with pytest.raises(ra.RaError):
        ex.map(d)
