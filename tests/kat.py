"""Known-answer tests ported from the reference's own test suite (values and shapes only; no code).

Each case is a function of an executor `ex` and is run twice, like the reference's TEST(plier) macro runs ply_ravel
and ply_fixed side by side (test/ply.cc:35-44): once on the CPU oracle (tests/test_oracle_golden.py, pins the
oracle) and once on the CUDA library (tests/test_gpu_kat.py, -m gpu). Citations are file:line in the reference.
"""
import numpy as np

from ra_ra_b200 import abi, ra

CASES = []


def case(f):
    CASES.append(f)
    return f


def arr(ex, values, dtype):
    return ex.from_numpy(np.array(values, dtype=dtype))


def zeros(ex, shape, dtype):
    return ex.from_numpy(np.zeros(shape, dtype=dtype))


def full(ex, shape, value, dtype):
    return ex.from_numpy(np.full(shape, value, dtype=dtype))


def eq(view_or_value, expected):
    got = view_or_value.numpy() if isinstance(view_or_value, ra.View) else np.asarray(view_or_value)
    expected = np.asarray(expected)
    assert got.shape == expected.shape, (got.shape, expected.shape)
    assert np.array_equal(got, expected), (got, expected)


def concrete(ex, e, shape, dtype):
    out = zeros(ex, shape, dtype)
    out.assign(e)
    return out


# ---------------- traversal / maps ----------------

@case
def ply_c_eq_a_minus_scalar(ex):
    """test/ply.cc:31-44: c = a - 3.0 on 3x2 -> {-3..2}."""
    a = arr(ex, np.arange(6).reshape(3, 2), np.float64)
    c = zeros(ex, (3, 2), np.float64)
    c.assign(a - 3.0)
    eq(c, np.arange(6).reshape(3, 2) - 3.0)


@case
def ply_int_neg_mul(ex):
    """test/ply.cc:51-61: c = -a -> {-1,-2,-3}; c = a*b -> {1,-4,9}."""
    A = arr(ex, [1, 2, 3], np.int32)
    B = arr(ex, [1, -2, 3], np.int32)
    C = zeros(ex, (3,), np.int32)
    C.assign(-A)
    eq(C, [-1, -2, -3])
    C.assign(A * B)
    eq(C, [1, -4, 9])


@case
def ply_rank3_vs_rank2_prefix(ex):
    """test/ply.cc:262-288: a(3,2,4) - b(3,2), 24-value check at :265-266."""
    a = arr(ex, np.arange(1, 25).reshape(3, 2, 4), np.float64)
    b = arr(ex, np.arange(1, 7).reshape(3, 2), np.float64)
    c = zeros(ex, (3, 2, 4), np.float64)
    check = [0, 1, 2, 3, 3, 4, 5, 6, 6, 7, 8, 9, 9, 10, 11, 12, 12, 13, 14, 15, 15, 16, 17, 18]
    c.assign(a - b)
    eq(c, np.array(check, dtype=np.float64).reshape(3, 2, 4))
    c.assign(0.)
    c.assign(-(b - a))
    eq(c, np.array(check, dtype=np.float64).reshape(3, 2, 4))


@case
def ply_empty_and_rank0(ex):
    """test/ply.cc:107-151,318-333: a zero-length axis means the op never runs; rank 0 runs once."""
    a = zeros(ex, (0, 11, 2), np.float64)
    c = full(ex, (1,), 99., np.float64)
    c0 = c(0)  # rank 0 view
    assert c0.rank == 0
    c0.assign(77.)
    eq(c, [77.])
    c0 += 1.
    eq(c, [78.])
    e = zeros(ex, (0, 11, 2), np.float64)
    e.assign(a + 1.)  # nothing to do, nothing to fail
    assert ra.sum(a) == 0. and ra.prod(a) == 1.  # test/reduction.cc:199-201


@case
def ply_tindex_and_scalar(ex):
    """test/ply.cc:245-259, test/frame-old.cc:28-58: _0/_1 index operands."""
    c = zeros(ex, (3, 2), np.int32)
    c.assign(ra._0 + ra._1)
    eq(c, [[0, 1], [1, 2], [2, 3]])
    c.assign(ra._0 * 2)
    eq(c, [[0, 0], [2, 2], [4, 4]])


@case
def mismatched_steps_transpose(ex):
    """test/ra-1.cc:183-203: c = a - transpose(b) with a(2,3), b(3,2) -> {0,-1,-2,2,1,0}; also through a transposed LHS."""
    a = arr(ex, np.arange(1, 7).reshape(2, 3), np.int32)
    b = arr(ex, np.arange(1, 7).reshape(3, 2), np.int32)
    c = zeros(ex, (2, 3), np.int32)
    c.assign(a - ra.transpose(b, (1, 0)))
    eq(c, [[0, -1, -2], [2, 1, 0]])
    c.assign(0)
    ra.transpose(c, (1, 0)).assign(ra.transpose(a, (1, 0)) - b)
    eq(c, [[0, -1, -2], [2, 1, 0]])


@case
def operators_sum_and_mask(ex):
    """test/operators.cc:122-137: a(3,2)+b(3) -> {11,12,23,40,35,36}; a==b mask; !(a!=b)."""
    a = arr(ex, [[1, 2], [3, 20], [5, 6]], np.int32)
    b = arr(ex, [10, 20, 30], np.int32)
    eq(concrete(ex, a + b, (3, 2), np.int32), [[11, 12], [23, 40], [35, 36]])
    mask = [[False, False], [False, True], [False, False]]
    eq(concrete(ex, a.eq(b), (3, 2), np.bool_), mask)
    eq(concrete(ex, ~(a.ne(b)), (3, 2), np.bool_), mask)


@case
def readme_first_example(ex):
    """examples/read-me.cc:26-37: B += A + C + std::vector{100., 200.}; B(all, iota(len/2, len/2)) *= -1."""
    v = [[1, 2, 3, 4], [5, 6, 7, 8]]
    A, B, Cc = arr(ex, v, np.float32), arr(ex, v, np.float32), arr(ex, v, np.float32)
    B += A + Cc + np.array([100., 200.])
    B(ra.all, ra.iota(ra.len / 2, ra.len / 2)).__imul__(-1)
    eq(B, np.array([[103, 106, -109, -112], [215, 218, -221, -224]], dtype=np.float32))


@case
def readme_vector_float(ex):
    """examples/read-me.cc:162-165: B += std::vector<float>{10, 20} -> {{11,12},{23,24}}."""
    B = arr(ex, [[1, 2], [3, 4]], np.float32)
    B += np.array([10, 20], dtype=np.float32)
    eq(B, np.array([[11, 12], [23, 24]], dtype=np.float32))


@case
def ra12_map_broadcast_reduce(ex):
    """test/ra-12.cc:30-43, examples/read-me.cc:82-91: C = A*B; D += A*B -> {-6, 6}."""
    A = arr(ex, [[1, 2, 3], [1, 2, 3]], np.float32)
    B = arr(ex, [-1, 1], np.float32)
    Cc = full(ex, (2, 3), 99., np.float32)
    Cc.assign(A * B)
    eq(Cc, np.array([[-1, -2, -3], [1, 2, 3]], dtype=np.float32))
    D = zeros(ex, (2,), np.float32)
    D += A * B
    eq(D, np.array([-6, 6], dtype=np.float32))


@case
def compat_vector_plus_big(ex):
    """test/compatibility.cc:22-40,61-66: foreign vector matched on axis 0, also through a transposed LHS."""
    a = zeros(ex, (2, 3), np.int32)
    a.assign(np.array([10, 20], dtype=np.int32) + ra._1)
    eq(a, [[10, 11, 12], [20, 21, 22]])
    b = zeros(ex, (3, 2), np.int32)
    ra.transpose(b).assign(np.array([10, 20], dtype=np.int32) + ra._1)
    eq(b, [[10, 20], [11, 21], [12, 22]])


@case
def frame_periodic_axes(ex):
    """test/frame-new.cc:146-172: b = a(all, insert<1>, iota(4, 0, 0)) gives a rank 4 match 2x3x4x4."""
    a = zeros(ex, (2, 3, 4), np.int32)
    a.assign((ra._0 + 1) * 100 + (ra._1 + 1) * 10 + (ra._2 + 1))
    b = a(ra.all, ra.insert(1), ra.iota(4, 0, 0))
    assert [d[0] for d in b.dims] == [2, abi.UNB, 4, 4]
    out = zeros(ex, (2, 3, 4, 4), np.int32)
    out.assign(a + b)
    an = a.numpy()
    expect = an[:, :, :, None] + an[:, None, 0, :][:, :, None, :].repeat(4, axis=2)
    # c(all, j) = a(all, iota(4,0,0)) for every j: c[i,j,k,l] = a[i,0,l]; a is matched on its prefix: a[i,j,k]
    expect = an[:, :, :, None] + an[:, 0, :][:, None, None, :]
    eq(out, expect)
    out.assign(b + a)  # order doesn't affect prefix matching
    eq(out, expect)


@case
def frame_outer_product(ex):
    """test/frame-new.cc:175-190: a(dots<2>, insert<1>) - b(insert<2>, dots<1>) is 4x3x5."""
    a = zeros(ex, (4, 3), np.int32)
    a.assign(10 * ra._1 + 100 * ra._0)
    b = zeros(ex, (5,), np.int32)
    b.assign(ra._0)
    out = zeros(ex, (4, 3, 5), np.int32)
    out.assign(a(ra.dots(2), ra.insert(1)) - b(ra.insert(2), ra.dots(1)))
    eq(out, a.numpy()[:, :, None] - b.numpy()[None, None, :])


@case
def operators_pluseq_chain(ex):
    """test/operators.cc:195-218: array+scalar and += chains."""
    a = arr(ex, [1, 2, 3], np.float64)
    a += 1
    a += a
    eq(a, [4., 6., 8.])
    a -= arr(ex, [1, 1, 1], np.int32)
    a /= 2.
    a *= a
    eq(a, [2.25, 6.25, 12.25])


# ---------------- views ----------------

@case
def viewops_reverse(ex):
    """test/view-ops.cc:19-52: reverse on each axis of 3x2x4; equivalence with iota(len, len-1, -1)."""
    a = arr(ex, np.arange(1, 25).reshape(3, 2, 4), np.float64)
    an = a.numpy()
    check0 = [17, 18, 19, 20, 21, 22, 23, 24, 9, 10, 11, 12, 13, 14, 15, 16, 1, 2, 3, 4, 5, 6, 7, 8]
    check1 = [5, 6, 7, 8, 1, 2, 3, 4, 13, 14, 15, 16, 9, 10, 11, 12, 21, 22, 23, 24, 17, 18, 19, 20]
    check2 = [4, 3, 2, 1, 8, 7, 6, 5, 12, 11, 10, 9, 16, 15, 14, 13, 20, 19, 18, 17, 24, 23, 22, 21]
    rev = ra.iota(ra.len, ra.len - 1, -1)
    for k, check, sub in ((0, check0, (rev,)), (1, check1, (ra.dots(1), rev)), (2, check2, (ra.dots(2), rev))):
        b = ra.reverse(a, k)
        eq(b, np.array(check, dtype=np.float64).reshape(3, 2, 4))
        c = a(*sub)
        assert c.ptr == b.ptr and c.dims == b.dims
        eq(b, np.flip(an, axis=k))


@case
def viewops_transpose(ex):
    """test/view-ops.cc:54-75: transpose of 3x2 iota(1..6) reads {1,3,5,2,4,6}; :121-138 diag."""
    a = arr(ex, np.arange(1, 7).reshape(3, 2), np.float64)
    eq(ra.transpose(a, (1, 0)), np.array([[1, 3, 5], [2, 4, 6]], dtype=np.float64))
    eq(ra.transpose(a), np.array([[1, 3, 5], [2, 4, 6]], dtype=np.float64))
    s = arr(ex, np.arange(9).reshape(3, 3), np.float64)
    eq(ra.diag(s), [0., 4., 8.])


@case
def fromb_offsets(ex):
    """test/fromb.cc:46-81: base pointer offsets 30, 0, 40, 34, 2, 12, 92 and values of beaten subscripts on a 10x10."""
    a = zeros(ex, (10, 10), np.int32)
    a.assign(ra._1 + 10 * ra._0)
    an = a.numpy()

    def off(b):
        return (b.ptr - a.ptr) // 4

    b = a(3, ra.all)
    assert off(b) == 30
    eq(b, an[3])
    b = a(ra.iota(4))
    assert off(b) == 0
    eq(b, an[:4])
    b = a(ra.iota(4, 4))
    assert off(b) == 40
    eq(b, an[4:8])
    b = a(3, ra.iota(5, 4))
    assert off(b) == 34
    eq(b, an[3, 4:9])
    b = a(ra.all, ra.iota(4, 2, 2))
    assert off(b) == 2
    eq(b, an[:, 2:10:2])
    b = a(ra.iota(3, 1, 2), ra.iota(2, 2, 3))
    assert off(b) == 12
    eq(b, an[1:7:2, 2:8:3])
    b = a(ra.iota(3, 9, -2), ra.iota(2, 2, 3))
    assert off(b) == 92
    eq(b, an[9:3:-2, 2:8:3])


@case
def fromb_len_and_insert(ex):
    """test/fromb.cc:131-162 (ra::len in subscripts), :195-211 (insert), :253-274 (a(iota(2,0,2), iota(2,0,2)))."""
    a = zeros(ex, (10, 10), np.int32)
    a.assign(ra._1 + 10 * ra._0)
    an = a.numpy()
    eq(a(ra.len - 1, ra.all), an[9])
    eq(a(ra.all, ra.len - 1), an[:, 9])
    eq(a(ra.iota(ra.len - 6, 3)), an[3:7])
    eq(a(ra.iota(2, 0, 2), ra.iota(2, 0, 2)), an[0:4:2, 0:4:2])
    b = a(ra.insert(1), ra.all)
    assert [d[0] for d in b.dims] == [abi.UNB, 10, 10] and b.dims[0][1] == 0


@case
def fromb_bounds_errors(ex):
    """ra/ply.hh:395,407 RA_CK; test/checks.cc:215-233 dynamic mismatch raises at ply."""
    a = zeros(ex, (10, 10), np.int32)
    for bad in ((10,), (-1,), (ra.iota(4, 8),), (ra.iota(3, 9, -5),)):
        try:
            a(*bad)
        except ra.RaError as e:
            assert e.status == abi.ERR_BOUNDS
        else:
            raise AssertionError("bounds not checked")
    b = zeros(ex, (3,), np.int32)
    try:
        a += b
    except ra.RaError as e:
        assert e.status == abi.ERR_SHAPE
    else:
        raise AssertionError("agreement not checked")
    c = zeros(ex, (3,), np.float64)
    try:
        c.assign(ra._1)  # test/checks.cc:159-193: unbounded
    except ra.RaError as e:
        assert e.status == abi.ERR_SHAPE
    else:
        raise AssertionError("unbounded expression not rejected")


@case
def iota_values(ex):
    """test/iota.cc:21-34 values; :36-49 transpose(b,{0,2,1}) = iota(3,1) ; :51-67 unbounded + bounded."""
    a = zeros(ex, (4,), np.int32)
    a.assign(ra.iota(4, 3))
    eq(a, [3, 4, 5, 6])
    a.assign(ra.iota(4, 3, 2))
    eq(a, [3, 5, 7, 9])
    a.assign(ra.iota(4) + ra.iota(4, 10))
    eq(a, [10, 12, 14, 16])
    b = zeros(ex, (3, 2, 4), np.int32)
    ra.transpose(b, (0, 2, 1)).assign(ra.iota(3, 1))
    eq(b, np.broadcast_to(np.array([1, 2, 3], dtype=np.int32)[:, None, None], (3, 2, 4)))
    c = zeros(ex, (3,), np.int32)
    c.assign(ra.iota() + ra.iota(3, 10))
    eq(c, [10, 12, 14])


# ---------------- where / pick ----------------

@case
def where_pick(ex):
    """test/where.cc:22-38: pick(p, a0, a1) -> {1,20,3}; where(p, a0, a1) -> {10,2,30}."""
    a0, a1 = arr(ex, [1, 2, 3], np.float64), arr(ex, [10, 20, 30], np.float64)
    p = arr(ex, [0, 1, 0], np.int32)
    a = zeros(ex, (3,), np.float64)
    a.assign(ra.pick(p, a0, a1))
    eq(a, [1., 20., 3.])
    a.assign(ra.where(p, a0, a1))
    eq(a, [10., 2., 30.])


@case
def where_rvalue(ex):
    """test/where.cc:98-109."""
    w = arr(ex, [True, False, False, True], np.bool_)
    eq(concrete(ex, ra.where(w, 1, 2), (4,), np.int32), [1, 2, 2, 1])
    v = arr(ex, [1, 2, 3, 4], np.int32)
    eq(concrete(ex, ra.where(ra.logical_and(ra._0 > 0, ra._0 < 3), v, 17), (4,), np.int32), [17, 2, 3, 17])
    eq(concrete(ex, ra.where(w, 2 * ra._0 + 1, 2 * ra._0), (4,), np.int32), [1, 2, 4, 7])
    a = zeros(ex, (4, 3), np.int32)
    a.assign(ra._0 - ra._1)
    eq(concrete(ex, ra.where(w, 99, a), (4, 3), np.int32), [[99, 99, 99], [1, 0, -1], [2, 1, 0], [99, 99, 99]])


@case
def where_nested(ex):
    """test/where.cc:110-122."""
    a = arr(ex, [-1, 0, 1], np.int32)
    eq(concrete(ex, ra.where(a >= 0, ra.where(a < 1, 77, 99), 44), (3,), np.int32), [44, 77, 99])


# ---------------- reductions ----------------

@case
def reductions_dot_sqrm(ex):
    """test/reduction.cc:85-109: dot({1,2},{3,4}) = 11, reduce_sqrm(a-b) = 8, norm2 = sqrt(8)."""
    a, b = arr(ex, [1, 2], np.float64), arr(ex, [3, 4], np.float64)
    assert ra.dot(a, b) == 11.
    assert ra.reduce_sqrm(a - b) == 8.
    assert ra.norm2(a - b) == np.sqrt(8.)


@case
def reductions_sum_prod_amax_amin(ex):
    """test/reduction.cc:123-138: sum/prod/amax/amin; NaN never enters amax/amin; empty defaults."""
    a = arr(ex, [4, 6, -3, 2], np.float64)
    assert ra.sum(a) == 9. and ra.prod(a) == -144. and ra.amax(a) == 6. and ra.amin(a) == -3.
    assert ra.amax(ra.abs(a)) == 6. and ra.amin(ra.abs(a)) == 2.
    q = full(ex, (3,), np.nan, np.float64)
    assert ra.amax(q) == -np.inf and ra.amin(q) == np.inf
    i = arr(ex, [4, 6, -3, 2], np.int32)
    assert ra.sum(i) == 9 and ra.amax(i) == 6 and ra.amin(i) == -3 and ra.prod(i) == -144
    e = zeros(ex, (0,), np.int32)
    assert ra.amax(e) == np.iinfo(np.int32).min and ra.amin(e) == np.iinfo(np.int32).max  # :224-239


@case
def reductions_sum_cols(ex):
    """test/reduction.cc:141-165: A(100,111) = _0-_1; C(100) += A and C = C + A."""
    A = zeros(ex, (100, 111), np.float64)
    A.assign(ra._0 - ra._1)
    B = A.numpy().sum(axis=1)
    Cc = zeros(ex, (100,), np.float64)
    Cc += A
    eq(Cc, B)
    Cc.assign(0.)
    Cc.assign(Cc + A)  # sequentially every step reads the C just written: a running sum (:160-164)
    eq(Cc, B)
    Cc.assign(A)  # without the self reference the last one wins per row (docs/ra-ra.texi:1986-2007)
    eq(Cc, A.numpy()[:, 110])


@case
def reductions_sum_rows(ex):
    """test/reduction.cc:182-222: sum rows of A(100,111) via transposed += and via leading insert."""
    A = zeros(ex, (100, 111), np.float64)
    A.assign(ra._0 - ra._1)
    B = A.numpy().sum(axis=0)
    Cc = zeros(ex, (111,), np.float64)
    Cc += ra.transpose(A)
    eq(Cc, B)
    Cc.assign(0.)
    cc = Cc(ra.insert(1))
    cc += A
    eq(Cc, B)


@case
def reductions_max_rows_cols(ex):
    """test/reduction.cc:224-239: m = max(m, c) with a broadcast destination is a running max: {1,3,2},{7,1,3} -> cols {7,3,3}, rows {3,7}."""
    c = arr(ex, [[1, 3, 2], [7, 1, 3]], np.int32)
    m = zeros(ex, (3,), np.int32)
    mm = m(ra.insert(1))
    mm.assign(ra.max(mm, c))
    eq(m, [7, 3, 3])
    r = zeros(ex, (2,), np.int32)
    r.assign(ra.max(r, c))
    eq(r, [3, 7])
    r.assign(100)
    r.assign(ra.min(r, c))
    eq(r, [1, 1])


@case
def gemv_gevm(ex):
    """test/reduction.cc:241-279: gemv/gevm -> {-20,-10,0}."""
    a = zeros(ex, (3, 4), np.float64)
    a.assign(ra._0 - ra._1)
    an = a.numpy()
    b = arr(ex, [1, 2, 3, 4], np.float64)
    c = zeros(ex, (3,), np.float64)
    ra.gemv(a, b, c)
    eq(c, an @ b.numpy())
    eq(c, [-20., -10., 0.])
    at = zeros(ex, (4, 3), np.float64)
    at.assign(ra._1 - ra._0)
    c.assign(0.)
    ra.gevm(b, at, c)
    eq(c, [-20., -10., 0.])


@case
def gemm_small(ex):
    """test/reduction.cc:280-333: 3x2 of 2s times 2x4 of 3s -> 12; c += form -> 13; 0x0 corner."""
    a, b = full(ex, (3, 2), 2., np.float64), full(ex, (2, 4), 3., np.float64)
    c = zeros(ex, (3, 4), np.float64)
    ra.gemm(a, b, c)
    eq(c, np.full((3, 4), 12.))
    c.assign(1.)
    ra.gemm(a, b, c)
    eq(c, np.full((3, 4), 13.))
    a0, b0, c0 = zeros(ex, (0, 2), np.float64), zeros(ex, (2, 0), np.float64), zeros(ex, (0, 0), np.float64)
    ra.gemm(a0, b0, c0)


@case
def gemm_integer_valued(ex):
    """bench/bench-gemm.cc:229-237: a = _0-_1, b = _1-2*_0, exact in any summation order."""
    m, k, n = 13, 17, 11
    a, b = zeros(ex, (m, k), np.float64), zeros(ex, (k, n), np.float64)
    a.assign(ra._0 - ra._1)
    b.assign(ra._1 - 2 * ra._0)
    c = zeros(ex, (m, n), np.float64)
    ra.gemm(a, b, c)
    eq(c, a.numpy() @ b.numpy())


# ---------------- stencils ----------------

@case
def stencil3_slices_vs_raw(ex):
    """bench/bench-stencil3.cc:55-64,130: f_slices equals f_raw; boundary never written; pointer swap semantics."""
    rng = np.random.default_rng(41)
    n = (9, 8, 11)
    A0 = rng.uniform(-1, 1, n).astype(np.float32)
    A, An = ex.from_numpy(A0), zeros(ex, n, np.float32)
    I, J, K = ra.iota(n[0] - 2, 1), ra.iota(n[1] - 2, 1), ra.iota(n[2] - 2, 1)
    ref, refn = A0.copy(), np.zeros(n, np.float32)
    for _ in range(4):  # "must be even bc of swap" test/wrank.cc:234
        An(I, J, K).assign(-6 * A(I, J, K) + A(I + 1, J, K) + A(I, J + 1, K) + A(I, J, K + 1)
                           + A(I - 1, J, K) + A(I, J - 1, K) + A(I, J, K - 1))
        A, An = An, A
        c = ref[1:-1, 1:-1, 1:-1]
        refn[1:-1, 1:-1, 1:-1] = ((((((np.float32(-6) * c + ref[2:, 1:-1, 1:-1]) + ref[1:-1, 2:, 1:-1]) + ref[1:-1, 1:-1, 2:])
                                    + ref[:-2, 1:-1, 1:-1]) + ref[1:-1, :-2, 1:-1]) + ref[1:-1, 1:-1, :-2])
        ref, refn = refn, ref
    eq(A, ref)
    eq(An, refn)


@case
def laplace2d_stiff_and_mass(ex):
    """examples/laplace2d.cc:36,48: 5-point stiffness and 7-point mass stencils on (N+1)^2 doubles."""
    rng = np.random.default_rng(7)
    N = 12
    v0 = rng.uniform(-1, 1, (N + 1, N + 1))
    v, w = ex.from_numpy(v0), zeros(ex, (N + 1, N + 1), np.float64)
    i, j = ra.iota(N - 1, 1), ra.iota(N - 1, 1)
    w(i, j).assign(4 * v(i, j) - v(i - 1, j) - v(i, j - 1) - v(i + 1, j) - v(i, j + 1))
    c = v0[1:-1, 1:-1]
    ref = np.zeros_like(v0)
    ref[1:-1, 1:-1] = (((4 * c - v0[:-2, 1:-1]) - v0[1:-1, :-2]) - v0[2:, 1:-1]) - v0[1:-1, 2:]
    eq(w, ref)
    h = 1. / N
    w(i, j).assign(ra.sqrm(h) * (v(i, j) / 2. + (v(i - 1, j) + v(i, j - 1) + v(i + 1, j) + v(i, j + 1) + v(i + 1, j + 1) + v(i - 1, j - 1)) / 12.))
    ref[1:-1, 1:-1] = (h * h) * (c / 2. + (((((v0[:-2, 1:-1] + v0[1:-1, :-2]) + v0[2:, 1:-1]) + v0[1:-1, 2:]) + v0[2:, 2:]) + v0[:-2, :-2]) / 12.)
    eq(w, ref)


@case
def laplace2d_cghs_end_to_end(ex):
    """examples/laplace2d.cc:60-91 + examples/cghs.hh:15-45: -laplace u = f on (N+1)^2 by conjugate gradients, every array op on
    the executor. Pinned by the reference: max |u - g| <= 0.00463 and <= 67 iterations at N = 50, eps = 1e-5."""
    N, EPS = 50, 1e-5
    h = 1. / N
    x = np.arange(N + 1) * h
    PI = np.pi
    g = np.zeros((N + 1, N + 1))
    f = np.zeros((N + 1, N + 1))
    g[1:-1, 1:-1] = np.sin(2 * PI * x[1:-1])[:, None] * np.sin(2 * PI * x[1:-1])[None, :]
    f[1:-1, 1:-1] = 8 * PI * PI * g[1:-1, 1:-1]
    v, w = ex.from_numpy(g), ex.from_numpy(f)
    u, b = zeros(ex, (N + 1, N + 1), np.float64), zeros(ex, (N + 1, N + 1), np.float64)
    i, j = ra.iota(N - 1, 1), ra.iota(N - 1, 1)

    def mult_stiff(vv, ww):  # laplace2d.cc:36
        ww(i, j).assign(4 * vv(i, j) - vv(i - 1, j) - vv(i, j - 1) - vv(i + 1, j) - vv(i, j + 1))

    def mult_mass(vv, ww):  # laplace2d.cc:48
        ww(i, j).assign(ra.sqrm(h) * (vv(i, j) / 2. + (vv(i - 1, j) + vv(i, j - 1) + vv(i + 1, j) + vv(i, j + 1)
                                                       + vv(i + 1, j + 1) + vv(i - 1, j - 1)) / 12.))
    mult_mass(w, b)
    work = zeros(ex, (3, N + 1, N + 1), np.float64)
    gg, r, p = work(0), work(1), work(2)
    err = EPS * EPS * ra.reduce_sqrm(b)
    mult_stiff(u, gg)
    gg.assign(b - gg)
    r.assign(gg)
    its = 0
    while ra.reduce_sqrm(gg) > err and its < 200:
        mult_stiff(r, p)
        rho = ra.reduce_sqrm(p)
        sig = ra.dot(r, p)
        tau = ra.dot(gg, r)
        t = tau / sig
        gam = (t * t * rho - tau) / tau
        u += float(t) * r
        gg -= float(t) * p
        r.assign(r * float(gam) + gg)
        its += 1
    mx = ra.amax(ra.abs(u - v))
    assert mx <= 0.00463, mx
    assert its <= 67, its


@case
def early_any_every_index(ex):
    """test/early.cc:17-82: any / every over rank 0, static and dynamic rank; index() incl. rank 2 ("only returns first index")."""
    a0 = arr(ex, 99, np.int32)                                     # rank 0 (:18-22)
    assert ra.any(a0.eq(99)) and ra.every(a0.eq(99))
    a = zeros(ex, (2, 3), np.int32)                                # a = _0*2 + _1*10 (:53-57): no odd element
    a.assign(ra.tindex(0, abi.I32) * 2 + ra.tindex(1, abi.I32) * 10)
    assert not ra.any((a % 2).ne(0)) and ra.every((a % 2).eq(0))
    b = arr(ex, [[1, 1, 1], [1, 0, 1]], np.int32)                  # :36-43, also through a transposed view
    assert not ra.every(b.eq(1)) and ra.any(b.eq(0))
    bt = ra.transpose(b, (1, 0))
    assert not ra.every(bt.eq(1)) and ra.any(bt.eq(0))
    c = zeros(ex, (9,), np.int32)                                  # a = iota(9)*2 (:59-63)
    c.assign(ra.iota(9) * 2)
    assert not ra.any((c % 2).ne(0)) and ra.every((c % 2).eq(0))
    d = arr(ex, [1, 3, -1, 9], np.int32)                           # :65-69
    assert ra.index(d < 0) == 2 and ra.index(d > 10) == -1
    e = arr(ex, [[1, 3], [-1, 9]], np.int32)                       # :77-82
    assert ra.index(e < 0) == 1 and ra.index(e > 10) == -1
    z = zeros(ex, (0,), np.int32)                                  # defaults of early() on an empty traversal (ra/ra.hh:282,288,297)
    assert not ra.any(z.eq(0)) and ra.every(z.eq(1)) and ra.index(z.eq(0)) == -1


@case
def lexical_compare_first_difference(ex):
    """ra/ra.hh:299-304: a < b at the first differing position in row-major order, False when equal."""
    a = arr(ex, [[1, 2, 3], [4, 5, 6]], np.int32)
    b = arr(ex, [[1, 2, 3], [4, 7, 0]], np.int32)
    assert ra.lexical_compare(a, b) and not ra.lexical_compare(b, a) and not ra.lexical_compare(a, a)
    c = arr(ex, [[1, 2, 4], [0, 0, 0]], np.int32)                  # differs earlier, the other way: later elements do not matter
    assert ra.lexical_compare(a, c) and not ra.lexical_compare(c, a)
    x, y = arr(ex, [0.5, 2.5], np.float64), arr(ex, [0.5, -1.], np.float64)
    assert not ra.lexical_compare(x, y) and ra.lexical_compare(y, x)


@case
def unbeaten_subscripts_gather(ex):
    """test/fromu.cc:23-31,46-67,70-83,107-121: a(i ...) with index arrays is the outer product of the subscripts applied to a
    (ra/ply.hh:461-499); scalars and iotas mix in; `len` resolves inside an unbeatable subscript."""
    a = zeros(ex, (10,), np.int32)
    a.assign(ra.tindex(0, abi.I32))
    eq(concrete(ex, a(ra.len - np.array([1, 2, 3, 4])), (4,), np.int32), [9, 8, 7, 6])                   # :25-26
    v = arr(ex, [7, 9, 3, 4], np.float64)
    eq(concrete(ex, v(np.array([3, 2, 0, 1])), (4,), np.float64), [4, 3, 7, 9])                           # :50
    eq(concrete(ex, v(arr(ex, [[3, 2], [0, 1]], np.int32)), (2, 2), np.float64), [[4, 3], [7, 9]])        # :63-64 rank-2 index
    m = arr(ex, [[1, 2], [3, 4]], np.float64)
    for i, j in (([0, 1], [0, 1]), ([0, 1], [1, 0]), ([1, 0], [0, 1]), ([1, 0], [1, 0])):                 # :75-82
        eq(concrete(ex, m(np.array(i), np.array(j)), (2, 2), np.float64), m.numpy()[np.ix_(i, j)])
    eq(concrete(ex, m(np.array([1, 0]), 1), (2,), np.float64), [4, 2])                                    # :111-118 mixed
    eq(concrete(ex, m(1, np.array([1, 0])), (2,), np.float64), [4, 3])
    eq(concrete(ex, m(np.array([0, 1]), 1), (2,), np.float64), [2, 4])
    # index array on one axis, the other kept whole (beaten `all`, implied): rows picked whole
    b = zeros(ex, (4, 3), np.int32)
    b.assign(ra.tindex(0, abi.I32) * 10 + ra.tindex(1, abi.I32))
    eq(concrete(ex, b(np.array([3, 0, 3])), (3, 3), np.int32), b.numpy()[[3, 0, 3]])
    eq(concrete(ex, b(ra.all, np.array([2, 0])), (4, 2), np.int32), b.numpy()[:, [2, 0]])
    eq(concrete(ex, b(ra.iota(2, 1), np.array([2, 0])), (2, 2), np.int32), b.numpy()[1:3][:, [2, 0]])
    # through a transposed / reversed source and inside a larger expression
    bt = ra.transpose(b, (1, 0))
    eq(concrete(ex, bt(np.array([2, 0]), np.array([3, 1, 0])) * 2 + 1, (2, 3), np.int32), b.numpy().T[np.ix_([2, 0], [3, 1, 0])] * 2 + 1)
    # at(a, i): index tuples (ra/ra.hh:228-239; examples/indirect.cc)
    ij = arr(ex, [[0, 1], [3, 2], [1, 0]], np.int32)
    eq(concrete(ex, ra.at(b, ij), (3,), np.int32), [1, 32, 10])


@case
def laplace2d_cghs_scalars_on_device(ex):
    """The same solve as laplace2d_cghs_end_to_end (examples/laplace2d.cc:60-91), through ra_ra_b200.cg.cghs: rho, sig, tau, t and
    gam never leave the executor's memory. Same pins: max |u - g| <= 0.00463, <= 67 iterations at N = 50, eps = 1e-5."""
    from ra_ra_b200 import cg
    N, EPS = 50, 1e-5
    h = 1. / N
    x = np.arange(N + 1) * h
    PI = np.pi
    g = np.zeros((N + 1, N + 1))
    f = np.zeros((N + 1, N + 1))
    g[1:-1, 1:-1] = np.sin(2 * PI * x[1:-1])[:, None] * np.sin(2 * PI * x[1:-1])[None, :]
    f[1:-1, 1:-1] = 8 * PI * PI * g[1:-1, 1:-1]
    v, w = ex.from_numpy(g), ex.from_numpy(f)
    u, b = zeros(ex, (N + 1, N + 1), np.float64), zeros(ex, (N + 1, N + 1), np.float64)
    i, j = ra.iota(N - 1, 1), ra.iota(N - 1, 1)

    def mult_stiff(vv, ww):  # laplace2d.cc:36
        ww(i, j).assign(4 * vv(i, j) - vv(i - 1, j) - vv(i, j - 1) - vv(i + 1, j) - vv(i, j + 1))
    b(i, j).assign(ra.sqrm(h) * (w(i, j) / 2. + (w(i - 1, j) + w(i, j - 1) + w(i + 1, j) + w(i, j + 1) + w(i + 1, j + 1) + w(i - 1, j - 1)) / 12.))
    work = zeros(ex, (3, N + 1, N + 1), np.float64)
    scal = zeros(ex, (8,), np.float64)
    its = cg.cghs(mult_stiff, b, u, work, EPS, scal, max_its=200)
    mx = ra.amax(ra.abs(u - v))
    assert mx <= 0.00463, mx
    assert its <= 67, its


@case
def unbeaten_subscripts_scatter(ex):
    """test/fromu.cc:28-31,53-61,84-92: a(i ...) with index arrays as an lvalue; the right hand side matches the subscripts' frame by
    prefix; repeated indices accumulate (sequential in the reference, atomic on the device: exact for integers)."""
    a = zeros(ex, (4,), np.float64)
    a(np.array([3, 2, 0, 1])).assign(np.array([9., 7., 1., 4.]))                       # :53-55
    eq(a, [1, 4, 7, 9])
    a.assign(0.)
    a(np.array([3, 2, 0, 1])).assign(77.)                                              # :59-61 rank extension of the right hand side
    eq(a, [77, 77, 77, 77])
    m = zeros(ex, (2, 2), np.float64)
    m(np.array([1, 0]), np.array([1, 0])).assign(arr(ex, [[9, 7], [1, 4]], np.float64))  # :84-86
    eq(m, [[4, 1], [7, 9]])
    g = m(np.array([1, 0]), np.array([1, 0]))
    g *= arr(ex, [[9, 7], [1, 4]], np.float64)                                         # :87-88
    eq(m, [[16, 1], [49, 81]])
    m(np.array([1, 0]), np.array([1, 0])).assign(np.array([9., 7.]))                   # :90-91 prefix matching: rows of the frame
    eq(m, [[7, 7], [9, 9]])
    b = zeros(ex, (10, 10), np.int32)
    b.assign(ra.tindex(0, abi.I32) - ra.tindex(1, abi.I32))
    g = b(ra.len - np.array([1, 3, 4]))                                                # :28-31 whole rows 9, 7, 6
    g += 100
    ref = np.arange(10)[:, None] - np.arange(10)[None, :]
    ref[[9, 7, 6]] += 100
    eq(b, ref)
    # repeated indices: a histogram (h(i) += 1), mixed types (float counts += int), and -= through a strided destination
    idx = arr(ex, [0, 3, 3, 1, 3, 0, 2, 3], np.int32)
    h = zeros(ex, (5,), np.int32)
    g = h(idx)
    g += 1
    eq(h, [2, 1, 1, 4, 0])
    hf = zeros(ex, (2, 5), np.float32)
    g = hf(1)(idx)
    g += arr(ex, [1, 2, 3, 4, 5, 6, 7, 8], np.int32)
    eq(hf, [[0, 0, 0, 0, 0], [7, 4, 7, 18, 0]])
    ht = ra.transpose(zeros(ex, (5, 2), np.float64), (1, 0))
    g = ht(0)(idx)
    g -= 0.5
    eq(ht, [[-1, -.5, -.5, -2, 0], [0, 0, 0, 0, 0]])
    # gather from one array, scatter into another, in one expression
    src = arr(ex, [10., 20., 30., 40.], np.float64)
    out = zeros(ex, (6,), np.float64)
    out(np.array([5, 0, 2])).assign(src(np.array([3, 3, 1])) * 2 + 1)
    eq(out, [81, 0, 41, 0, 0, 81])


@case
def stencil_view_sumprod_form(ex):
    """ra::stencil (ra/arrays.hh:615-630) and the f_sumprod form of bench/bench-stencil3.cc:94-105,130: Anext(I,J,K) += Astencil * mask
    with the mask on the neighbourhood axes gives the same sweep as the slice expression of :55-64 (the bench checks its variants
    against each other); also the 2-D case of bench-stencil2.cc and test/wrank.cc:276."""
    rng = np.random.default_rng(3)
    n = (7, 6, 9)
    a0 = rng.integers(-8, 9, n).astype(np.float64)
    A, An1, An2 = ex.from_numpy(a0), zeros(ex, n, np.float64), zeros(ex, n, np.float64)
    I, J, K = (ra.iota(m - 2, 1) for m in n)
    An1(I, J, K).assign(-6 * A(I, J, K) + A(I + 1, J, K) + A(I, J + 1, K) + A(I, J, K + 1) + A(I - 1, J, K) + A(I, J - 1, K) + A(I, J, K - 1))
    mask = np.zeros((3, 3, 3))
    mask[1, 1, 1] = -6
    for d in ((0, 1, 1), (2, 1, 1), (1, 0, 1), (1, 2, 1), (1, 1, 0), (1, 1, 2)):
        mask[d] = 1
    S = ra.stencil(A, 1, 1)
    assert list(S.shape) == [5, 4, 7, 3, 3, 3], S.shape
    eq(concrete(ex, S(0, 0, 0, ra.dots(3)), (3, 3, 3), np.float64), a0[:3, :3, :3])    # the neighbourhood of cell (1,1,1) (:141-142)
    g = An2(I, J, K)
    g += S * ex.from_numpy(mask)(ra.insert(3))
    eq(An2, An1.numpy())
    b0 = rng.integers(-8, 9, (6, 8)).astype(np.float64)
    B, Bn = ex.from_numpy(b0), zeros(ex, (6, 8), np.float64)
    g = Bn(ra.iota(4, 1), ra.iota(6, 1))
    g += ra.stencil(B, 1, 1) * arr(ex, [[0, 1, 0], [1, -4, 1], [0, 1, 0]], np.float64)(ra.insert(2))
    ref = np.zeros((6, 8))
    ref[1:-1, 1:-1] = -4 * b0[1:-1, 1:-1] + b0[2:, 1:-1] + b0[1:-1, 2:] + b0[:-2, 1:-1] + b0[1:-1, :-2]
    eq(Bn, ref)


@case
def maxwell_vector_potential(ex):
    """examples/maxwell.cc:29-88: rank 5/6 slice assignments with `len`, shifted iotas, += on slices, transpose<0,1,2,3,5,4>,
    the diagonal reduce divA += transpose<0,1,2,3,4,4>(DA). Pinned by the reference: amax(divA) == 0 and
    F(19,0,0,0,2,1) == 0.3039588939177449 (exact where cos() is the host libm's; 1e-12 on the device's cos)."""
    H, PI, delta = ra.all, np.pi, 1.
    o, n, m, l = 20, 20, 2, 2
    A = zeros(ex, (o, n, m, l, 4), np.float64)
    DA, F = zeros(ex, (o, n, m, l, 4, 4), np.float64), zeros(ex, (o, n, m, l, 4, 4), np.float64)
    divA = zeros(ex, (o, n, m, l), np.float64)
    X, Y = zeros(ex, (n, m, l, 4), np.float64), zeros(ex, (n, m, l, 4), np.float64)
    A(0, H, H, H, 2).assign(-ra.cos(ra.iota(n) * (2 * PI / n)) / (2 * PI / n))                     # :38
    A(1, H, H, H, 2).assign(-ra.cos((ra.iota(n) - delta) * (2 * PI / n)) / (2 * PI / n))           # :39
    L = ra.len
    for t in range(1, o - 1):                                                                     # :43-60
        X(ra.iota(L - 1)).assign(A(t, ra.iota(L - 1, 1)))
        X(L - 1).assign(A(t, 0))
        g = X(H, ra.iota(L - 1)); g += A(t, H, ra.iota(L - 1, 1))
        g = X(H, L - 1); g += A(t, H, 0)
        g = X(H, H, ra.iota(L - 1)); g += A(t, H, H, ra.iota(L - 1, 1))
        g = X(H, H, L - 1); g += A(t, H, H, 0)
        Y(ra.iota(L - 1, 1)).assign(A(t, ra.iota(L - 1)))
        Y(0).assign(A(t, L - 1))
        g = Y(H, ra.iota(L - 1, 1)); g += A(t, H, ra.iota(L - 1))
        g = Y(H, 0); g += A(t, H, L - 1)
        g = Y(H, H, ra.iota(L - 1, 1)); g += A(t, H, H, ra.iota(L - 1))
        g = Y(H, H, 0); g += A(t, H, H, L - 1)
        A(t + 1).assign(X + Y - A(t - 1) - 4 * A(t))
    for k, factor in ((0, +1 / (2 * delta)), (1, -1 / (2 * delta)), (2, -1 / (2 * delta)), (3, -1 / (2 * delta))):   # diff, :64-71
        if DA.len(k) >= 2:
            hk, h4k = ra.dots(k), ra.dots(4 - k)
            DA(hk, ra.iota(L - 2, 1), h4k, k).assign(A(hk, ra.iota(L - 2, 2)) - A(hk, ra.iota(L - 2, 0)))
            DA(hk, 0, h4k, k).assign(A(hk, 1) - A(hk, L - 1))
            DA(hk, L - 1, h4k, k).assign(A(hk, 0) - A(hk, L - 2))
            g = DA(ra.dots(5), k); g *= factor
    F.assign(ra.transpose(DA, (0, 1, 2, 3, 5, 4)) - DA)                                           # :80
    divA += ra.transpose(DA, (0, 1, 2, 3, 4, 4))                                                   # :81 prefix matching reduces the last axis
    assert ra.amax(divA) == 0.                                                                     # :83 Lorentz gauge
    f = float(F(19, 0, 0, 0, 2, 1).numpy())
    if ex.name == "cuda":
        assert abs(f - 0.3039588939177449) <= 1e-12, f
    else:
        assert f == 0.3039588939177449, repr(f)                                                    # :88
    Ey, Bz = F(ra.dots(1), H, 0, 0, 2, 0), F(ra.dots(1), H, 0, 0, 2, 1)                            # :90-92 fields stay within [-1, 1]
    assert ra.amin(Ey) >= -1 and ra.amax(Ey) <= 1 and ra.amin(Bz) >= -1 and ra.amax(Bz) <= 1


@case
def checkply_reverse_transpose_combinations(ex):
    """test/ra-1.cc:42-62,225-264: B -= A then B += A; B -= A for every reverse / transpose combination of source and destination."""
    def check(mkA, mkB, shape):
        a0 = np.arange(1, 1 + int(np.prod(shape)), dtype=np.int32).reshape(shape)
        A0, B0 = ex.from_numpy(a0), ex.from_numpy(a0)
        A, B = mkA(A0), mkB(B0)
        c = B.numpy().copy()
        an = A.numpy()
        B -= A
        eq(B, c - an)
        B += A
        B -= A
        eq(B, c - an)
    rev = lambda *ks: (lambda v: [v := ra.reverse(v, k) for k in ks][-1])  # noqa: E731
    ident = lambda v: v  # noqa: E731
    for mkA, mkB in ((ident, ident), (rev(0), ident), (ident, rev(0)), (rev(0), rev(0)), (rev(1), ident), (ident, rev(1)), (rev(1), rev(1)),
                     (ident, rev(0, 1)), (rev(0, 1), ident), (rev(0, 1), rev(0, 1))):
        check(mkA, mkB, (2, 3))
    tr = lambda v: ra.transpose(v, (1, 0))  # noqa: E731
    trr = lambda v: ra.reverse(ra.reverse(ra.transpose(v, (1, 0)), 1), 0)  # noqa: E731
    for mkA, mkB in ((tr, ident), (ident, tr), (trr, ident), (ident, trr)):
        check(mkA, mkB, (2, 2))


@case
def agreement_examples(ex):
    """examples/agreement.cc:22-36,41-55,62-66,70-84: X = A+B-C with shapes [3 4 5], [3 4], [3]; B*7, B*C, B*B; i-j needs a shape;
    A = b (prefix) against A = c(insert<1>) (skip an axis)."""
    A, B, C = full(ex, (3, 4, 5), 1., np.float32), full(ex, (3, 4), 2., np.float32), full(ex, (3,), 3., np.float32)
    X = full(ex, (3, 4, 5), 99., np.float32)
    X.assign(A + B - C)
    eq(X, np.zeros((3, 4, 5), np.float32))
    eq(concrete(ex, B * 7., (3, 4), np.float32), np.full((3, 4), 14., np.float32))
    c2 = arr(ex, [1, 2, 3], np.float32)
    eq(concrete(ex, B * c2, (3, 4), np.float32), np.array([[2.] * 4, [4.] * 4, [6.] * 4], np.float32))
    eq(concrete(ex, B * B, (3, 4), np.float32), np.full((3, 4), 4., np.float32))
    ij = zeros(ex, (3, 4), np.float32)
    ij.assign(ra.tindex(0) - ra.tindex(1))
    eq(ij, (np.arange(3)[:, None] - np.arange(4)[None, :]).astype(np.float32))
    A2, b, c = zeros(ex, (3, 4), np.float32), arr(ex, [0, 1, 2], np.float32), arr(ex, [0, 1, 2, 3], np.float32)
    A2.assign(b)
    eq(A2, np.repeat(np.arange(3, dtype=np.float32)[:, None], 4, 1))
    A2.assign(c(ra.insert(1)))
    eq(A2, np.repeat(np.arange(4, dtype=np.float32)[None, :], 3, 0))
    try:
        ra.sum(ra.tindex(0) - ra.tindex(1))                          # no shape anywhere: "the following would be invalid" (:60)
        raise AssertionError("an expression without a shape must not evaluate")
    except ra.RaError as e:
        assert e.status == abi.ERR_SHAPE


@case
def frame_old_full_expression_only(ex):
    """test/frame-old.cc:73-86 (c = b - a with a(3,2), b(3)) and :107-119: agreement is checked on the WHOLE expression: b - _1 has no
    shape of its own along axis 1 but the destination gives it one."""
    a = ex.from_numpy(np.arange(1, 7, dtype=np.float64).reshape(3, 2))
    b = arr(ex, [1, 2, 3], np.float64)
    c = zeros(ex, (3, 2), np.float64)
    c.assign(a - b)
    eq(c, [[0, 1], [1, 2], [2, 3]])
    a2, b2 = zeros(ex, (2, 2), np.float64), arr(ex, [1., 2.], np.float64)
    a2.assign(b2 - ra.tindex(1))
    eq(a2, [[1, 0], [2, 1]])


@case
def where_as_lvalue(ex):
    """test/where.cc:148-157 [raop01]: where(cond, a, b) = 7 writes a where cond holds and b elsewhere; += likewise."""
    a, b = zeros(ex, (4,), np.int32), zeros(ex, (4,), np.int32)
    i = ra.tindex(0, abi.I32)
    ra.where(ra.logical_and(i > 0, i < 3), a, b).assign(7)
    eq(a, [0, 7, 7, 0])
    eq(b, [7, 0, 0, 7])
    g = ra.where(ra.logical_or(i <= 0, i >= 3), a, b)
    g += 2
    eq(a, [2, 7, 7, 2])
    eq(b, [7, 2, 2, 7])
    m = arr(ex, [True, False], np.bool_)                              # :163-166 where with index operands as branches
    eq(concrete(ex, ra.where(m, 3 * ra.tindex(0, abi.I32), 2 * ra.tindex(0, abi.I32)), (2,), np.int32), [0, 2])


@case
def unbounded_expression_raises(ex):
    """test/checks.cc:159-193: a view with an inserted (unbounded) axis cannot be traversed on its own."""
    z = zeros(ex, (10,), np.float64)
    w = z(ra.insert(1), ra.all)
    try:
        w *= -1
        raise AssertionError("w *= -1 over an unbounded axis must raise")
    except ra.RaError as e:
        assert e.status == abi.ERR_SHAPE and "unbounded" in str(e)
    eq(z, np.zeros(10))


@case
def stencil2d_four_steps(ex):
    """test/wrank.cc:230-280 / bench/bench-stencil2.cc:49-58,124: four sweeps of the 2-D 5-point stencil with buffer swap ("ts must be
    even bc of swap"), slices against plain loops."""
    n = 9
    rng = np.random.default_rng(8)
    a0 = rng.integers(-4, 5, (n, n)).astype(np.float64)
    A, An = ex.from_numpy(a0), zeros(ex, (n, n), np.float64)
    I, J = ra.iota(n - 2, 1), ra.iota(n - 2, 1)
    ref, refn = a0.copy(), np.zeros((n, n))
    for _ in range(4):
        An(I, J).assign(-4 * A(I, J) + A(I + 1, J) + A(I, J + 1) + A(I - 1, J) + A(I, J - 1))
        A, An = An, A
        refn[1:-1, 1:-1] = -4 * ref[1:-1, 1:-1] + ref[2:, 1:-1] + ref[1:-1, 2:] + ref[:-2, 1:-1] + ref[1:-1, :-2]
        ref, refn = refn, ref
    eq(A, ref)
    eq(An, refn)


@case
def bench_self_checks(ex):
    """The exactness checks the reference's benches make on themselves: bench/bench-gemv.cc:87-92 (a = _0-_1, b = 1-2*_0), bench-sum-rows.cc:26-34
    and bench-sum-cols.cc:23-31 (a = _0-_1: every spelling gives the same integers), bench-reduce-sqrm.cc:31 and bench-dot.cc (constant fills)."""
    m, n = 37, 53
    a = zeros(ex, (m, n), np.float64)
    a.assign(ra.tindex(0) - ra.tindex(1))
    an = np.arange(m)[:, None] - np.arange(n)[None, :] + 0.
    b = zeros(ex, (n,), np.float64)
    b.assign(1 - 2 * ra.tindex(0))
    c = zeros(ex, (m,), np.float64)
    ra.gemv(a, b, c)
    eq(c, an @ (1. - 2 * np.arange(n)))
    bt = zeros(ex, (m,), np.float64)
    bt.assign(1 - 2 * ra.tindex(0))
    c2 = zeros(ex, (n,), np.float64)
    ra.gevm(bt, a, c2)
    eq(c2, (1. - 2 * np.arange(m)) @ an)
    rows, cols = zeros(ex, (n,), np.float64), zeros(ex, (m,), np.float64)
    g = rows(ra.insert(1))
    g += a
    cols += a
    eq(rows, an.sum(0))
    eq(cols, an.sum(1))
    N = 1000
    x = full(ex, (N, 4), 1.5, np.float32)
    y = full(ex, (N, 4), 2.0, np.float32)
    assert ra.reduce_sqrm(x) == np.float32(2.25 * 4 * N)
    assert ra.dot(x, y) == np.float32(3.0 * 4 * N)
    assert ra.reduce_sqrm(x - y) == np.float32(0.25 * 4 * N)


def _laplace3d_fem(ex, n):
    """examples/laplace3d.cc:24-176: trilinear FEM for -laplace u = f on [-1,1]^3; the operator is gather -> 8x8 element matrix
    -> scatter-add over all elements (mult0, :60-68), written here as ONE fused fold (the gather feeds the element gemv) and ONE
    scatter per product, then CG with device-resident scalars."""
    from ra_ra_b200 import cg
    d00, d16, d23 = 0., -1. / 6., 2. / 3.
    SM = np.array([[d23, d00, d16, d00, d00, d16, d16, d16], [d00, d23, d00, d16, d16, d00, d16, d16],
                   [d16, d00, d23, d00, d16, d16, d00, d16], [d00, d16, d00, d23, d16, d16, d16, d00],
                   [d00, d16, d16, d16, d23, d00, d16, d00], [d16, d00, d16, d16, d00, d23, d00, d16],
                   [d16, d16, d00, d16, d16, d00, d23, d00], [d16, d16, d16, d00, d00, d16, d00, d23]])
    d827, d427, d227, d127 = 8. / 27., 4. / 27., 2. / 27., 1. / 27.
    MM = np.array([[d827, d427, d227, d427, d427, d227, d127, d227], [d427, d827, d427, d227, d227, d427, d227, d127],
                   [d227, d427, d827, d427, d127, d227, d427, d227], [d427, d227, d427, d827, d227, d127, d227, d427],
                   [d427, d227, d127, d227, d827, d427, d227, d427], [d227, d427, d227, d127, d427, d827, d427, d227],
                   [d127, d227, d427, d227, d227, d427, d827, d427], [d227, d127, d227, d427, d427, d227, d427, d827]])
    ISF = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]])
    h, PI = 1. / n, np.pi
    pos = lambda m, i, j, k: m * (m * i + j) + k  # noqa: E731
    gi = np.arange(n + 1)
    X, Y, Z = np.meshgrid(-1. + 2. * gi * h, -1. + 2. * gi * h, -1. + 2. * gi * h, indexing="ij")   # V[pos(n+1,i,j,k)] (:113-116)
    ei, ej, ek = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    EV = np.stack([pos(n + 1, ei + ISF[v][0], ej + ISF[v][1], ek + ISF[v][2]).ravel() for v in range(8)], 1).astype(np.int32)  # :119-122
    bi, bj = np.meshgrid(gi, gi, indexing="ij")
    Bnd = np.concatenate([pos(n + 1, 0, bi, bj).ravel(), pos(n + 1, n, bi, bj).ravel(), pos(n + 1, bi, 0, bj).ravel(),
                          pos(n + 1, bi, n, bj).ravel(), pos(n + 1, bi, bj, 0).ravel(), pos(n + 1, bi, bj, n).ravel()]).astype(np.int32)  # :125-135
    nv, ne = (n + 1) ** 3, n ** 3
    ev, bnd = ex.from_numpy(EV), ex.from_numpy(Bnd)
    T = zeros(ex, (ne, 8), np.float64)

    def make_mult(M, scale):
        Md = ex.from_numpy(M)

        def mult(v, w):                                              # mult0 (:60-68)
            T.assign(0.)
            t = T
            t += Md(ra.insert(1)) * v(ev(ra.all, ra.insert(1)))      # T(e,a) = sum_b M(a,b) v[EV(e,b)]: gather fused into the fold
            w.assign(0.)
            g = w(ev)
            g += T                                                   # w(E.v) += M . v(E.v): scatter-add, vertices shared by 8 elements
            w(bnd).assign(0.)                                        # set boundary
            w *= scale                                               # scale factor
        return mult
    fx = .25 * 3. * PI * PI * np.cos(.5 * PI * X) * np.cos(.5 * PI * Y) * np.cos(.5 * PI * Z)
    gx = np.cos(.5 * PI * X) * np.cos(.5 * PI * Y) * np.cos(.5 * PI * Z)
    aux, c = ex.from_numpy(fx.ravel()), ex.from_numpy(gx.ravel())
    b, x = zeros(ex, (nv,), np.float64), zeros(ex, (nv,), np.float64)
    make_mult(MM, h * h * h)(aux, b)                                 # integral ~ product with the mass matrix (:147-149)
    work, scal = zeros(ex, (3, nv), np.float64), zeros(ex, (8,), np.float64)
    its = cg.cghs(make_mult(SM, h), b, x, work, 1e-5, scal, max_its=50)
    return float(ra.amax(ra.abs(c - x))), its


@case
def laplace3d_fem_gather_scatter(ex):
    """examples/laplace3d.cc:175-176 pins err <= 0.00033 and its <= 1 at n = 50 (cos cos cos is an eigenfunction of the discrete operator,
    so CG converges at once). n = 50 on the oracle and on the device; n = 12 on the (slow) planner simulator, error bound scaled by (50/12)^2."""
    n = 12 if ex.name == "sim" else 50
    err, its = _laplace3d_fem(ex, n)
    assert err <= 0.00033 * (50. / n) ** 2, err
    assert its <= 1, its    # the reference's own pin, on every executor (device: 5 runs at n = 50, its = 1 each time, profiles/r02_fem_its.log)


@case
def reference_reductions_refmin_refmax(ex):
    """test/reduction.cc:404-424: refmin / refmax return a reference to the FIRST extreme element; assigning through it changes the array."""
    A = zeros(ex, (2, 3), np.float64)
    A.assign(ra.tindex(1) - ra.tindex(0))                 # {{0, 1, 2}, {-1, 0, 1}}
    B = A.numpy().copy()
    mn = ra.refmin(A)
    eq(mn, -1.)
    mn.assign(-99.)
    B[1, 0] = -99
    eq(A, B)
    mx = ra.refmax(A)                                     # (the reference spells this refmin(A, std::greater<double>()))
    eq(mx, 2.)
    mx.assign(0.)
    B[0, 2] = 0
    eq(A, B)
    my = ra.refmax(A)                                     # two elements equal 1: the first one in row-major order, A(0, 1)
    eq(my, 1.)
    my.assign(77.)
    B[0, 1] = 77
    eq(A, B)


@case
def named_functions_opcode_table(ex):
    """The rest of the named functions of ra/ra.hh:210,221-222 and `<=>` (ra/ra.hh:190). Known answers: test/operators.cc:58 (odd of
    {1, 2, 3, -1} -> {true, false, true, true}), :65-68 (`<=>` of {3, 4, 5} with {2., 4., 6.} -> "+0-"), test/test.cc:86-88 (isfinite of
    3., +inf, nan), test/reexported.cc:25-26 (lerp(4., 1., 0.5) = 2.5); clamp / lerp / classification exact, libm functions to 4 ulp."""
    eq(concrete(ex, ra.odd(arr(ex, [1, 2, 3, -1], np.int32)), (4,), np.bool_), [True, False, True, True])
    c3 = concrete(ex, ra.cmp3(arr(ex, [3, 4, 5], np.int32), arr(ex, [2., 4., 6.], np.float64)), (3,), np.int32)
    eq(c3, [1, 0, -1])
    assert "".join("+0-"[1 - int(z)] for z in c3.numpy()) == "+0-"
    eq(concrete(ex, ra.cmp3(arr(ex, [1., np.nan], np.float64), arr(ex, [np.nan, 2.], np.float64)), (2,), np.int32), [2, 2])  # unordered
    x = arr(ex, [3., np.inf, np.nan, -np.inf, -0.], np.float64)
    eq(concrete(ex, ra.isfinite(x), (5,), np.bool_), [True, False, False, False, True])
    eq(concrete(ex, ra.isnan(x), (5,), np.bool_), [False, False, True, False, False])
    eq(concrete(ex, ra.isinf(x), (5,), np.bool_), [False, True, False, True, False])
    xf = arr(ex, [3., np.inf, np.nan], np.float32)
    eq(concrete(ex, ra.isfinite(xf), (3,), np.bool_), [True, False, False])
    eq(concrete(ex, ra.lerp(arr(ex, [4., 4., 4., -1.], np.float64), 1., arr(ex, [0.5, 0., 1., 2.], np.float64)), (4,), np.float64), [2.5, 4., 1., 3.])
    eq(concrete(ex, ra.lerp(arr(ex, [1., 2.], np.float32), arr(ex, [3., 2.], np.float32), arr(ex, [0.25, 7.], np.float32)), (2,), np.float32),
       np.array([1.5, 2.], np.float32))
    eq(concrete(ex, ra.clamp(arr(ex, [-5, 0, 3, 9], np.int32), -1, 4), (4,), np.int32), [-1, 0, 3, 4])
    eq(concrete(ex, ra.clamp(arr(ex, [-5., 0.5, 9.], np.float64), arr(ex, [0., 0., 0.], np.float64), 1.), (3,), np.float64), [0., 0.5, 1.])
    eq(concrete(ex, ra.rel_error(arr(ex, [1., 0., 2.], np.float64), arr(ex, [1., 0., 1.], np.float64)), (3,), np.float64), [0., 0., 2. * 1. / 3.])
    eq(concrete(ex, ra.rel_error(arr(ex, [1., 3.], np.float32), arr(ex, [3., 1.], np.float32)), (2,), np.float32), np.array([1., 1.], np.float32))
    v = np.array([-0.75, -0.1, 0.0, 0.3, 0.9])
    for dt, ulp in ((np.float64, 4 * np.finfo(np.float64).eps), (np.float32, 4 * np.finfo(np.float32).eps)):
        a = arr(ex, v, dt)
        for fn, ref in ((ra.tan, np.tan), (ra.sinh, np.sinh), (ra.cosh, np.cosh), (ra.tanh, np.tanh), (ra.asin, np.arcsin), (ra.acos, np.arccos),
                        (ra.atan, np.arctan), (ra.expm1, np.expm1), (ra.log1p, np.log1p)):
            got = concrete(ex, fn(a), (5,), dt).numpy()
            want = ref(v.astype(dt).astype(np.float64))
            assert np.all(np.abs(got - want) <= ulp * np.maximum(1.0, np.abs(want))), (fn.__name__, got, want)
        got = concrete(ex, ra.log10(a + 2), (5,), dt).numpy()
        assert np.all(np.abs(got - np.log10(v.astype(dt).astype(np.float64) + 2)) <= ulp)
