"""CUDA path vs the CPU oracle on the same seeded inputs: bit-exact for maps, strided copies, integer folds and stencils;
stated tolerance vs an fp64 reference for float reductions."""
import numpy as np
import pytest

from oracle_exec import OracleExecutor
from randexpr import DT, SHAPES, random_expr, random_scatter, random_view
from ra_ra_b200 import abi, ra

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def exc():
    from ra_ra_b200 import cuda
    return cuda.CudaExecutor()


def snapshot(ex, dst):
    if ex.name == "cuda":
        return dst.owner.cpu().numpy().copy()
    return np.array(dst.owner, copy=True)


def case_random_maps(seed, executors):
    """The outputs of every executor for the four assignment forms of one seed (None where the expression is too long for a descriptor)."""
    rng = np.random.default_rng(1000 + seed)
    shape = SHAPES[seed % len(SHAPES)]
    res = []
    for assign in ("assign", "__iadd__", "__isub__", "__imul__"):
        state = rng.bit_generator.state
        outs = []
        for ex in executors:
            rng.bit_generator.state = state
            dst, _ = random_view(rng, ex, shape, DT[rng.integers(0, 4)])
            e = random_expr(rng, ex, shape, 2)
            try:
                getattr(dst, assign)(e)
            except ra.RaError as err:
                assert err.status == abi.ERR_UNSUPPORTED and ("too long" in str(err) or "too many" in str(err)), err
                outs = None
                break
            outs.append(snapshot(ex, dst))
        res.append((assign, outs))
    return res


@pytest.mark.parametrize("seed", range(60))
def test_random_maps_cuda_vs_oracle(seed, exc):
    for assign, outs in case_random_maps(seed, (OracleExecutor(), exc)):
        if outs is not None:
            assert np.array_equal(outs[0], outs[1], equal_nan=True), (seed, assign)


def case_random_scatter(seed, executors):
    outs = []
    for ex in executors:
        rng = np.random.default_rng(9000 + seed)
        outs.append(snapshot(ex, random_scatter(rng, ex)))
    return outs


@pytest.mark.parametrize("seed", range(40))
def test_random_scatter_cuda_vs_oracle(seed, exc):
    """a(i) OP= expr with index arrays: atomic read-modify-write on the device, sequential in the oracle; the generator keeps the
    result independent of the order (distinct indices, or integer += / -= with repeats)."""
    outs = case_random_scatter(seed, (OracleExecutor(), exc))
    assert outs[0].dtype == outs[1].dtype and np.array_equal(outs[0], outs[1]), seed


def test_scatter_histogram_large(exc):
    """2^22 increments into 1000 bins (heavy contention on the compare-and-swap loop): counts are exact."""
    import torch
    n, bins = 1 << 22, 1000
    idx_t = torch.randint(0, bins, (n,), device="cuda", dtype=torch.int32)
    h_t = torch.zeros(bins, device="cuda", dtype=torch.int32)
    g = exc.wrap(h_t)(exc.wrap(idx_t))
    g += 1
    assert exc.last_kernel().startswith("ply_map")
    assert torch.equal(h_t.long(), torch.bincount(idx_t.long(), minlength=bins))


def case_random_folds(seed, executors):
    rng = np.random.default_rng(5000 + seed)
    shape = [s for s in SHAPES if len(s) >= 2][seed % 8]
    res = []
    for variant in range(3):
        state = rng.bit_generator.state
        outs = []
        for ex in executors:
            rng.bit_generator.state = state
            rank = len(shape)
            if variant == 0:
                r = int(rng.integers(0, rank))
                dst, _ = random_view(rng, ex, shape[:r], DT[rng.integers(0, 4)], True)
            elif variant == 1:
                r = int(rng.integers(1, rank))
                dst, _ = random_view(rng, ex, shape[r:], DT[rng.integers(0, 4)], True)
                dst = dst(ra.insert(r))
            else:
                k = int(rng.integers(0, rank))
                dst, _ = random_view(rng, ex, shape[:k] + shape[k + 1:], DT[rng.integers(0, 4)], True)
                dst = dst(ra.dots(k), ra.insert(1))
            a, _ = random_view(rng, ex, shape, DT[rng.integers(0, 4)], True)
            b, _ = random_view(rng, ex, shape[:int(rng.integers(0, rank + 1))], DT[rng.integers(0, 4)], True)
            op = ["__iadd__", "__isub__", "assign"][int(rng.integers(0, 3))]
            getattr(dst, op)(a * b + 1)
            outs.append(snapshot(ex, dst))
        res.append(outs)
    return res


@pytest.mark.parametrize("seed", range(40))
def test_random_folds_cuda_vs_oracle(seed, exc):
    for variant, outs in enumerate(case_random_folds(seed, (OracleExecutor(), exc))):
        assert np.array_equal(outs[0], outs[1]), (seed, variant)


@pytest.mark.parametrize("n", [1, 3, 4, 5, 255, 256, 257, 1023, 4096, 65537, 1 << 20, (1 << 22) + 3])
@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32])
def test_contiguous_map_sizes(n, dt, exc):
    """B += A*C and B += A*C + scalar over every tail / alignment case of the vectorised fast path."""
    rng = np.random.default_rng(n)
    a, b, c = (rng.integers(-50, 50, n).astype(dt) for _ in range(3))
    outs = []
    for ex in (OracleExecutor(), exc):
        A, B, Cc = ex.from_numpy(a), ex.from_numpy(b), ex.from_numpy(c)
        B += A * Cc
        B += A * Cc + 3
        Bs = B(ra.iota(max(n - 1, 0), min(1, n - 1) if n > 1 else 0)) if n > 1 else B  # misaligned start
        Bs -= 1
        outs.append(snapshot(ex, B))
    assert np.array_equal(outs[0], outs[1])


@pytest.mark.parametrize("shape", [(1000,), (1 << 20,), (333, 777), (17, 19, 23), (4099, 513)])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_float_reductions_tolerance(shape, dt, exc):
    """sum / reduce_sqrm / dot in unspecified order (docs/ra-ra.texi:3102): relative error vs an fp64 reference.
    Tolerance: 8 eps(dt) relative to sum|x| (pairwise-tree bound, far below the sequential CPU fold's n*eps)."""
    rng = np.random.default_rng(11)
    x = rng.uniform(-1, 1, shape).astype(dt)
    y = rng.uniform(-1, 1, shape).astype(dt)
    X, Y = exc.from_numpy(x), exc.from_numpy(y)
    eps = np.finfo(dt).eps
    x64, y64 = x.astype(np.float64), y.astype(np.float64)
    for got, ref, scale in ((ra.sum(X), x64.sum(), np.abs(x64).sum()),
                            (ra.reduce_sqrm(X), (x64 * x64).sum(), (x64 * x64).sum()),
                            (ra.dot(X, Y), (x64 * y64).sum(), np.abs(x64 * y64).sum()),
                            (ra.reduce_sqrm(X - Y), ((x64 - y64) ** 2).sum(), ((x64 - y64) ** 2).sum())):
        assert abs(float(got) - ref) <= 8 * eps * scale, (got, ref, scale)
    assert ra.amax(X) == x.max() and ra.amin(X) == x.min()
    # deterministic: identical bits run to run
    assert ra.sum(X) == ra.sum(X) and ra.dot(X, Y) == ra.dot(X, Y)


@pytest.mark.parametrize("shape", [(300, 1024), (1024, 300), (64, 64), (5, 100000), (100000, 5), (2000, 33)])
def test_row_col_sums(shape, exc):
    """bench/bench-sum-rows.cc:26-34, bench/bench-sum-cols.cc:23-31: a = _0 - _1 is integer valued, so exact."""
    m, n = shape
    a = exc.empty(shape, abi.F32)
    a.assign(ra._0 - ra._1)
    an = (np.arange(m)[:, None] - np.arange(n)[None, :]).astype(np.float64)
    cols = exc.from_numpy(np.zeros(m, dtype=np.float64))
    cols += a
    assert np.array_equal(cols.numpy(), an.sum(axis=1))
    rows = exc.from_numpy(np.zeros(n, dtype=np.float64))
    rr = rows(ra.insert(1))
    rr += a
    assert np.array_equal(rows.numpy(), an.sum(axis=0))
    rows.assign(0.)
    rows += ra.transpose(a)
    assert np.array_equal(rows.numpy(), an.sum(axis=0))


@pytest.mark.parametrize("n", [(40, 36, 50), (64, 64, 64), (9, 130, 7)])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_stencil3d_bitexact(n, dt, exc):
    """bench/bench-stencil3.cc:55-64 through racu_map, against the oracle's f_raw loops: bit-exact, boundaries untouched."""
    import ctypes as C
    from oracle_exec import lib as olib
    rng = np.random.default_rng(41)
    A0 = rng.uniform(-1, 1, n).astype(dt)
    A, An = exc.from_numpy(A0), exc.from_numpy(np.zeros(n, dtype=dt))
    I, J, K = ra.iota(n[0] - 2, 1), ra.iota(n[1] - 2, 1), ra.iota(n[2] - 2, 1)
    ref, refn = A0.copy(), np.zeros(n, dtype=dt)
    for _ in range(4):
        An(I, J, K).assign(-6 * A(I, J, K) + A(I + 1, J, K) + A(I, J + 1, K) + A(I, J, K + 1)
                           + A(I - 1, J, K) + A(I, J - 1, K) + A(I, J, K - 1))
        A, An = An, A
        s = abi.StencilDesc()
        s.kind, s.dtype, s.rank = abi.STENCIL_3D7, ra.NP2RACU[np.dtype(dt)], 3
        s.src, s.dst = ref.ctypes.data, refn.ctypes.data
        for k in range(3):
            s.len[k] = n[k]
            s.src_stride[k] = s.dst_stride[k] = ref.strides[k] // ref.itemsize
        assert olib().oracle_stencil_raw(C.byref(s)) == 0
        ref, refn = refn, ref
    assert np.array_equal(A.numpy(), ref) and np.array_equal(An.numpy(), refn)


def test_static_specialisations_are_taken_and_bitexact(exc):
    """Flat single-type programs run on the compile-time specialised kernels (static_kernels.cu); same bits as the oracle."""
    rng = np.random.default_rng(77)
    for dt in (np.float32, np.float64):
        for n in (4096, 100003, (1 << 20) + 2):
            a, b, c = (rng.uniform(-1, 1, n).astype(dt) for _ in range(3))
            t = dt(0.37)
            outs = []
            for ex in (OracleExecutor(), exc):
                A, B, Cc = ex.from_numpy(a), ex.from_numpy(b), ex.from_numpy(c)
                kernels = []
                B += A * Cc
                kernels.append(ex.last_kernel() if ex.name == "cuda" else None)
                B -= t * A          # cghs: g -= t*p
                kernels.append(ex.last_kernel() if ex.name == "cuda" else None)
                B.assign(B * t + Cc)  # cghs: r = r*gam + g
                kernels.append(ex.last_kernel() if ex.name == "cuda" else None)
                Cc.assign(A)
                Cc /= B
                outs.append((snapshot(ex, B), snapshot(ex, Cc)))
                if ex.name == "cuda":
                    assert all(k == "sflat_map" for k in kernels), kernels
                    ra.sum(A)
                    assert ex.last_kernel() == "sflat_fold_inner"
            assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    # row / column sums of a contiguous matrix take the static fold kernels too
    M = exc.from_numpy(rng.uniform(-1, 1, (512, 1024)).astype(np.float32))
    c = exc.from_numpy(np.zeros(512, dtype=np.float32))
    c += M
    assert exc.last_kernel() == "sflat_fold_inner"
    r = exc.from_numpy(np.zeros(1024, dtype=np.float32))
    rr = r(ra.insert(1))
    rr += M
    assert exc.last_kernel() == "sflat_fold_outer"
    Mn = M.numpy().astype(np.float64)
    assert np.allclose(c.numpy(), Mn.sum(axis=1), rtol=1e-5, atol=1e-4) and np.allclose(r.numpy(), Mn.sum(axis=0), rtol=1e-5, atol=1e-4)


def _stencil_desc(kind, dt, src_ptr, dst_ptr, shape, strides, lo=0, hi=0):
    s = abi.StencilDesc()
    s.kind, s.dtype, s.rank = kind, ra.NP2RACU[np.dtype(dt)], len(shape)
    s.src, s.dst = src_ptr, dst_ptr
    for k in range(len(shape)):
        s.len[k] = shape[k]
        s.src_stride[k] = s.dst_stride[k] = strides[k]
    s.lo0, s.hi0 = lo, hi
    return s


@pytest.mark.parametrize("shape", [(132, 260), (64, 64), (33, 1028), (7, 9), (50, 51)])
@pytest.mark.parametrize("kind", [abi.STENCIL_2D5, abi.STENCIL_2D5_STIFF])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_stencil2d_entry_bitexact(shape, kind, dt, exc):
    """racu_stencil (bench-stencil2.cc:53-55, laplace2d.cc:36) vs the oracle's raw loops; TMA kernel when the shape allows,
    fused ply kernels otherwise - both must be bit exact and leave the boundary alone."""
    import ctypes as C
    import torch
    from oracle_exec import lib as olib
    from ra_ra_b200 import cuda
    rng = np.random.default_rng(5)
    a = rng.uniform(-1, 1, shape).astype(dt)
    ref = np.full(shape, 7, dtype=dt)
    st = [a.strides[k] // a.itemsize for k in range(2)]
    assert olib().oracle_stencil_raw(C.byref(_stencil_desc(kind, dt, a.ctypes.data, ref.ctypes.data, shape, st))) == 0
    A = torch.from_numpy(a).cuda()
    out = torch.full(shape, 7, dtype=A.dtype, device="cuda")
    cuda.check(cuda.lib.racu_stencil(C.byref(_stencil_desc(kind, dt, A.data_ptr(), out.data_ptr(), shape, st)), cuda.stream_ptr()))
    assert np.array_equal(out.cpu().numpy(), ref), cuda.lib.racu_last_kernel()


@pytest.mark.parametrize("shape,lo,hi", [((70, 40, 132), 0, 0), ((70, 40, 132), 1, 2), ((70, 40, 132), 30, 69), ((20, 19, 21), 3, 11), ((130, 16, 256), 0, 0)])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_stencil3d_entry_planes(shape, lo, hi, dt, exc):
    """racu_stencil restricted to planes [lo0, hi0) of axis 0 (what the slab-sharded caller uses for boundary-first sweeps)."""
    import ctypes as C
    import torch
    from oracle_exec import lib as olib
    from ra_ra_b200 import cuda
    rng = np.random.default_rng(6)
    a = rng.uniform(-1, 1, shape).astype(dt)
    ref = np.full(shape, -3, dtype=dt)
    st = [a.strides[k] // a.itemsize for k in range(3)]
    assert olib().oracle_stencil_raw(C.byref(_stencil_desc(abi.STENCIL_3D7, dt, a.ctypes.data, ref.ctypes.data, shape, st, lo, hi))) == 0
    A = torch.from_numpy(a).cuda()
    out = torch.full(shape, -3, dtype=A.dtype, device="cuda")
    cuda.check(cuda.lib.racu_stencil(C.byref(_stencil_desc(abi.STENCIL_3D7, dt, A.data_ptr(), out.data_ptr(), shape, st, lo, hi)), cuda.stream_ptr()))
    assert np.array_equal(out.cpu().numpy(), ref), cuda.lib.racu_last_kernel()
    if shape[2] % 4 == 0:
        assert cuda.lib.racu_last_kernel() == b"stencil_tma<3D7>"


def test_stencil3d_large_roundtrip_property(exc):
    """Full-size property (no oracle at 512^3): the sweep is linear, so S(a + 2b) == S(a) + 2 S(b) up to rounding, boundaries
    are untouched, and a constant field has a zero interior after one sweep."""
    import ctypes as C
    import torch
    from ra_ra_b200 import cuda
    n = 512
    g = torch.Generator(device="cuda")
    g.manual_seed(41)
    a = torch.rand(n, n, n, device="cuda", generator=g)
    out = torch.full_like(a, 5.0)
    st = list(a.stride())
    cuda.check(cuda.lib.racu_stencil(C.byref(_stencil_desc(abi.STENCIL_3D7, np.float32, a.data_ptr(), out.data_ptr(), (n, n, n), st)), cuda.stream_ptr()))
    assert cuda.lib.racu_last_kernel() == b"stencil_tma<3D7>"
    assert bool((out[0] == 5).all()) and bool((out[-1] == 5).all()) and bool((out[:, 0] == 5).all()) and bool((out[:, :, -1] == 5).all())
    c = a[1:-1, 1:-1, 1:-1]
    ref = ((((((-6 * c + a[2:, 1:-1, 1:-1]) + a[1:-1, 2:, 1:-1]) + a[1:-1, 1:-1, 2:]) + a[:-2, 1:-1, 1:-1]) + a[1:-1, :-2, 1:-1]) + a[1:-1, 1:-1, :-2])
    assert torch.equal(out[1:-1, 1:-1, 1:-1], ref)
    a.fill_(3.25)
    cuda.check(cuda.lib.racu_stencil(C.byref(_stencil_desc(abi.STENCIL_3D7, np.float32, a.data_ptr(), out.data_ptr(), (n, n, n), st)), cuda.stream_ptr()))
    assert float(out[1:-1, 1:-1, 1:-1].abs().max()) == 0.0


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32])
def test_static_nd_maps_bitexact(dt, exc):
    """Row-structured maps on the specialised kernel: prefix-broadcast operand (examples/read-me.cc:163), reversed inner
    axis (ra/arrays.hh:408-428), strided outer axes (ra/ply.hh:399-415), rank 3 and 4."""
    rng = np.random.default_rng(123)
    shape = (6, 10, 24)
    a, b, c = (rng.integers(-9, 9, shape).astype(dt) for _ in range(3))
    v = rng.integers(-9, 9, shape[0]).astype(dt)
    w = rng.integers(-9, 9, shape[:2]).astype(dt)
    big = rng.integers(-9, 9, (7, 12, 21, 32)).astype(dt)
    outs = []
    for ex in (OracleExecutor(), exc):
        A, B, Cc, Vv, Ww, G = (ex.from_numpy(x) for x in (a, b, c, v, w, big))
        ks = []

        def k():
            ks.append(ex.last_kernel() if ex.name == "cuda" else None)
        B += A * Cc + Vv
        k()                      # B += A*C + v: examples/read-me.cc:32
        B += Vv
        k()                      # y += row value
        B -= Ww
        k()
        B.assign(ra.reverse(A, 2))
        k()
        Cc.assign(ra.reverse(ra.reverse(A, 2), 0) * B)
        k()
        H = ex.from_numpy(np.zeros((7, 4, 21, 16), dtype=dt))
        H.assign(G(ra.all, ra.iota(4, 1, 3), ra.all, ra.iota(16, 8)))   # strided outer, offset inner (still aligned)
        k()
        H += G(ra.all, ra.iota(4, 11, -3), ra.all, ra.iota(16, 31, -1))  # reversed outer and inner
        k()
        outs.append((snapshot(ex, B), snapshot(ex, Cc), snapshot(ex, H)))
        if ex.name == "cuda":
            assert ks == ["sflat_map"] * 7, ks
    for x, y in zip(*outs):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("mnk", [(128, 128, 32), (256, 384, 64), (128, 256, 96), (384, 128, 160)])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_gemm_tensor_core_path(mnk, dt, exc):
    """gemm(a, b, c) (ra/ra.hh:403-412) on the MMA kernels: exact for the integer-valued inputs of bench/bench-gemm.cc:229-237
    (any summation order), and within a stated bound of an fp64 reference for random inputs. `c +=` semantics, views with
    leading dimensions larger than the extent."""
    m, n, k = mnk
    # (i) integer valued: bit exact vs the oracle's k-ascending loops
    a = exc.empty((m, k), ra.NP2RACU[np.dtype(dt)])
    b = exc.empty((k, n), ra.NP2RACU[np.dtype(dt)])
    a.assign((ra._0 - ra._1) % 7)
    b.assign((ra._1 - 2 * ra._0) % 5)
    c = exc.from_numpy(np.ones((m, n), dtype=dt))
    ra.gemm(a, b, c)
    assert exc.last_kernel() in ("dgemm_dmma884", "sgemm_3xtf32_tcgen05"), exc.last_kernel()
    ref = np.ones((m, n)) + a.numpy().astype(np.float64) @ b.numpy().astype(np.float64)
    assert np.array_equal(c.numpy().astype(np.float64), ref)
    # (ii) random, through views of larger buffers (lda/ldb/ldc > extents)
    rng = np.random.default_rng(51)
    A = rng.uniform(-1, 1, (m, k + 8)).astype(dt)
    B = rng.uniform(-1, 1, (k, n + 16)).astype(dt)
    C0 = rng.uniform(-1, 1, (m, n + 4)).astype(dt)
    Av, Bv, Cv = exc.from_numpy(A)(ra.all, ra.iota(k)), exc.from_numpy(B)(ra.all, ra.iota(n)), exc.from_numpy(C0)
    Cs = Cv(ra.all, ra.iota(n))
    ra.gemm(Av, Bv, Cs)
    assert exc.last_kernel() in ("dgemm_dmma884", "sgemm_3xtf32_tcgen05")
    got = Cv.numpy().astype(np.float64)
    A64, B64 = A[:, :k].astype(np.float64), B[:, :n].astype(np.float64)
    ref = C0.astype(np.float64)
    ref[:, :n] += A64 @ B64
    scale = np.abs(A64) @ np.abs(B64)
    err = np.abs(got[:, :n] - ref[:, :n])
    if dt == np.float64:
        bound = 2 * k * np.finfo(np.float64).eps * scale + 1e-300
    else:  # 3xTF32: fp32 accumulation (k * 2^-24) plus the dropped lo*lo terms and tf32 rounding of lo (~2^-21 per product)
        bound = (k * 2.0 ** -24 + 2.0 ** -20) * scale + 1e-30
    assert (err <= bound).all(), (err.max(), (err / bound).max())
    assert np.array_equal(got[:, n:], C0[:, n:].astype(np.float64))  # columns outside the view untouched


def test_gemv_gevm_static_rows(exc):
    """gemv: c += a * b(insert<1>) (ra/ra.hh:418) is a row-wise fold on the specialised kernel; bench/bench-gemv.cc:87-92 values are exact."""
    m, n = 256, 512
    a = exc.empty((m, n), abi.F64)
    a.assign(ra._0 - ra._1)
    b = exc.empty((n,), abi.F64)
    b.assign(1 - 2 * ra._0)
    c = exc.from_numpy(np.zeros(m))
    ra.gemv(a, b, c)
    assert exc.last_kernel() == "sflat_fold_inner", exc.last_kernel()
    assert np.array_equal(c.numpy(), a.numpy() @ b.numpy())
    c2 = exc.from_numpy(np.zeros(n))
    x = exc.empty((m,), abi.F64)
    x.assign(1 + ra._0)
    ra.gevm(x, a, c2)
    assert exc.last_kernel() == "sflat_fold_outer", exc.last_kernel()
    assert np.array_equal(c2.numpy(), x.numpy() @ a.numpy())


def test_inplace_stencil_is_rejected_not_raced(exc):
    """A(I,J,K) = -6*A(I,J,K) + A(I+1,J,K) ... written onto A itself: the sequential reference has a defined (order dependent) result;
    a parallel sweep would race. Every entry (racu_map by pattern, racu_stencil, racu_stencil_push) must reject it like the planner."""
    import ctypes as C
    import torch
    from ra_ra_b200 import cuda
    n = (36, 40, 132)
    a_t = torch.rand(n, device="cuda")
    before = a_t.clone()
    A = exc.wrap(a_t)
    I, J, K = (ra.iota(n[k] - 2, 1) for k in range(3))
    with pytest.raises(ra.RaError) as e:
        A(I, J, K).assign(-6 * A(I, J, K) + A(I + 1, J, K) + A(I, J + 1, K) + A(I, J, K + 1) + A(I - 1, J, K) + A(I, J - 1, K) + A(I, J, K - 1))
    assert e.value.status == abi.ERR_UNSUPPORTED and "overlaps" in str(e.value)
    st = list(a_t.stride())
    s = _stencil_desc(abi.STENCIL_3D7, np.float32, a_t.data_ptr(), a_t.data_ptr(), n, st)
    assert cuda.lib.racu_stencil(C.byref(s), cuda.stream_ptr()) == abi.ERR_UNSUPPORTED
    h = abi.HaloPush()
    flags = torch.zeros(8, dtype=torch.int32, device="cuda")
    h.flags = flags.data_ptr()
    assert cuda.lib.racu_stencil_push(C.byref(s), C.byref(h), cuda.stream_ptr()) == abi.ERR_UNSUPPORTED
    # shifted by one plane: still overlapping
    s2 = _stencil_desc(abi.STENCIL_3D7, np.float32, a_t.data_ptr(), a_t.data_ptr() + 4 * st[0], (n[0] - 1, n[1], n[2]), st)
    assert cuda.lib.racu_stencil(C.byref(s2), cuda.stream_ptr()) == abi.ERR_UNSUPPORTED
    torch.cuda.synchronize()
    assert torch.equal(a_t, before)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_gemm_with_aliased_destination_is_rejected(dt, exc):
    import ctypes as C
    import torch
    from ra_ra_b200 import cuda
    n = 256
    a = torch.ones(n, n, device="cuda", dtype=torch.float32 if dt == np.float32 else torch.float64)
    b = torch.ones_like(a)
    for cc in (a, b):
        g = abi.GemmDesc()
        g.dtype, g.M, g.N, g.K = (abi.F32 if dt == np.float32 else abi.F64), n, n, n
        g.a, g.lda, g.b, g.ldb, g.c, g.ldc = a.data_ptr(), n, b.data_ptr(), n, cc.data_ptr(), n
        assert cuda.lib.racu_gemm(C.byref(g), cuda.stream_ptr()) == abi.ERR_UNSUPPORTED
    torch.cuda.synchronize()
    assert bool((a == 1).all()) and bool((b == 1).all())


@pytest.mark.parametrize("seed", range(8))
def test_narrowing_fold_cuda_vs_oracle(seed, exc):
    """Integer destination folded over a floating point right hand side: truncation at every step (ra/expr.hh:50-53)."""
    outs = []
    for ex in (OracleExecutor(), exc):
        rng = np.random.default_rng(31000 + seed)
        m, n, k = (int(x) for x in rng.integers(2, 6, 3))
        a = ex.from_numpy(rng.uniform(-3, 3, (m, n, k)).astype([np.float32, np.float64][seed % 2]))
        c = ex.from_numpy(rng.integers(-5, 5, (m,) if seed % 3 == 0 else (m, n)).astype([np.int32, np.int64][(seed // 2) % 2]))
        getattr(c, ["__iadd__", "__isub__", "__imul__"][seed % 3])(a)
        outs.append(snapshot(ex, c))
    assert outs[0].dtype == outs[1].dtype and np.array_equal(outs[0], outs[1]), seed


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [130, 257])
def test_shifted_slice_expressions_run_specialised(dt, n, exc):
    """User-written stencils over slices shifted by one element - rows that are not 16-byte aligned and whose length is not a whole
    number of vectors - run on the specialised map kernels (element-wise loads for the unaligned roles, partial vectors at the row ends),
    not on the interpreter, and equal the oracle bit for bit: the 2-D 7-point mass stencil of examples/laplace2d.cc:48 (with its /2. and
    /12. association), the 1-D 3-point stencil of bench/bench-stencil1.cc:45, and an unaligned flat copy."""
    rng = np.random.default_rng(77)
    v0 = rng.uniform(-1, 1, (n, n)).astype(dt)
    a0 = rng.uniform(-1, 1, n * n).astype(dt)
    h = 1.0 / n
    outs = []
    for ex in (OracleExecutor(), exc):
        v, w = ex.from_numpy(v0), ex.from_numpy(np.zeros_like(v0))
        i, j = ra.iota(n - 2, 1), ra.iota(n - 2, 1)
        w(i, j).assign(ra.sqrm(h) * (v(i, j) / 2. + (v(i - 1, j) + v(i, j - 1) + v(i + 1, j) + v(i, j + 1) + v(i + 1, j + 1) + v(i - 1, j - 1)) / 12.))
        k1 = ex.last_kernel() if ex.name == "cuda" else None
        A, An = ex.from_numpy(a0), ex.from_numpy(np.zeros_like(a0))
        I = ra.iota(n * n - 2, 1)
        An(I).assign(-2 * A(I) + A(I + 1) + A(I - 1))
        k2 = ex.last_kernel() if ex.name == "cuda" else None
        B = ex.from_numpy(np.zeros_like(a0))
        B(ra.iota(n * n - 3, 2)).assign(A(ra.iota(n * n - 3, 1)) * 2)
        k3 = ex.last_kernel() if ex.name == "cuda" else None
        outs.append((snapshot(ex, w), snapshot(ex, An), snapshot(ex, B)))
        if ex.name == "cuda":
            assert k1 in ("jit_map", "stencil_tma<2D7_MASS>") and k2 == "stencil_1d3" and k3 in ("jit_map", "sflat_map"), (k1, k2, k3)
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(67, 131), (9, 1031), (130, 2051), (33, 4099), (3, 3), (4, 70), (21, 13, 259)])
def test_user_stencils_on_odd_rows(dt, shape, exc):
    """Slice expressions over rows TMA cannot describe (odd pitch) on the lane-wise map kernel: 5-point stiffness and 7-point mass stencils
    (examples/laplace2d.cc:36,48), the 3-D 7-point sweep (bench/bench-stencil3.cc:59-61), a stencil over TWO arrays with `+=` onto a
    destination that is also an operand: all bit-exact against the oracle, boundaries untouched."""
    rng = np.random.default_rng(93)
    v0 = rng.uniform(-1, 1, shape).astype(dt)
    u0 = rng.uniform(-1, 1, shape).astype(dt)
    outs = []
    for ex in (OracleExecutor(), exc):
        v, u = ex.from_numpy(v0), ex.from_numpy(u0)
        res, kern = [], []
        if len(shape) == 2:
            i, j = ra.iota(shape[0] - 2, 1), ra.iota(shape[1] - 2, 1)
            w = ex.from_numpy(np.full(shape, 3, dtype=dt))
            w(i, j).assign(4 * v(i, j) - v(i - 1, j) - v(i, j - 1) - v(i + 1, j) - v(i, j + 1))
            res.append(snapshot(ex, w)); kern.append(ex.last_kernel() if ex.name == "cuda" else None)
            w = ex.from_numpy(np.full(shape, 3, dtype=dt))
            w(i, j).assign(0.3 * (v(i, j) / 2. + (v(i - 1, j) + v(i, j - 1) + v(i + 1, j) + v(i, j + 1) + v(i + 1, j + 1) + v(i - 1, j - 1)) / 12.))
            res.append(snapshot(ex, w)); kern.append(ex.last_kernel() if ex.name == "cuda" else None)
            w = ex.from_numpy(np.full(shape, 3, dtype=dt))
            w(i, j).__iadd__(v(i, j + 1) * u(i + 1, j) - v(i, j - 1) * u(i - 1, j) + u(i, j) * w(i, j) + v(i - 1, j))   # two classes + the destination as an operand
            res.append(snapshot(ex, w)); kern.append(ex.last_kernel() if ex.name == "cuda" else None)
        else:
            I, J, K = (ra.iota(n - 2, 1) for n in shape)
            w = ex.from_numpy(np.full(shape, 3, dtype=dt))
            w(I, J, K).assign(-6 * v(I, J, K) + v(I + 1, J, K) + v(I, J + 1, K) + v(I, J, K + 1) + v(I - 1, J, K) + v(I, J - 1, K) + v(I, J, K - 1))
            res.append(snapshot(ex, w)); kern.append(ex.last_kernel() if ex.name == "cuda" else None)
        if ex.name == "cuda" and min(shape) >= 4:
            assert all(k in ("jit_map", "sflat_map") or k.startswith("stencil_tma") for k in kern), kern   # never the interpreter
        outs.append(res)
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(66, 132), (35, 260), (200, 1028), (3, 4), (18, 136)])
def test_mass_stencil_on_the_tma_kernel(dt, shape, exc):
    """examples/laplace2d.cc:48 on arrays whose rows are 16-byte multiples: recognised in racu_map and run by the TMA stencil kernel
    (7 points incl. the two diagonal neighbours, /2. and /12. as IEEE divisions, for float arrays the six summed in float and the rest in
    double, like C++ promotes the expression). Bit-exact against the oracle; the boundary of w is not written."""
    rng = np.random.default_rng(91)
    v0 = rng.uniform(-1, 1, shape).astype(dt)
    h = 0.37
    outs = []
    for ex in (OracleExecutor(), exc):
        v, w = ex.from_numpy(v0), ex.from_numpy(np.full(shape, 3, dtype=dt))
        i, j = ra.iota(shape[0] - 2, 1), ra.iota(shape[1] - 2, 1)
        w(i, j).assign(ra.sqrm(h) * (v(i, j) / 2. + (v(i - 1, j) + v(i, j - 1) + v(i + 1, j) + v(i, j + 1) + v(i + 1, j + 1) + v(i - 1, j - 1)) / 12.))
        if ex.name == "cuda":
            assert ex.last_kernel() == "stencil_tma<2D7_MASS>", ex.last_kernel()
        w2 = ex.from_numpy(np.full(shape, 3, dtype=dt))
        w2(i, j).assign(0.25 * (v(i, j) / 3. + (v(i - 1, j) + v(i, j - 1) + v(i + 1, j) + v(i, j + 1) + v(i + 1, j + 1) + v(i - 1, j - 1)) / 7.))  # other scalars, no sqrm
        w3 = ex.from_numpy(np.full(shape, 3, dtype=dt))
        w3(i, j).assign(0.25 * (v(i, j) / 3. + (v(i - 1, j) + v(i, j - 1) + v(i + 1, j) + v(i, j + 1) + v(i + 1, j + 1) + v(i - 1, j + 1)) / 7.))  # NOT the mass stencil: last neighbour differs
        if ex.name == "cuda":
            assert ex.last_kernel() != "stencil_tma<2D7_MASS>"
        outs.append((snapshot(ex, w), snapshot(ex, w2), snapshot(ex, w3)))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [3, 4, 5, 130, 1027, 4096, 70001, (1 << 20) + 2])
def test_stencil_1d3(dt, n, exc):
    """bench/bench-stencil1.cc:45 as the slice expression and through racu_stencil: the shuffle kernel, bit-exact, ends untouched."""
    import ctypes as C
    import torch
    from oracle_exec import lib as olib
    from ra_ra_b200 import cuda
    rng = np.random.default_rng(92)
    a0 = rng.uniform(-1, 1, n).astype(dt)
    outs = []
    for ex in (OracleExecutor(), exc):
        A, An = ex.from_numpy(a0), ex.from_numpy(np.full(n, 9, dtype=dt))
        I = ra.iota(n - 2, 1)
        An(I).assign(-2 * A(I) + A(I + 1) + A(I - 1))
        if ex.name == "cuda":
            assert ex.last_kernel() == "stencil_1d3", ex.last_kernel()
        outs.append(snapshot(ex, An))
    assert np.array_equal(outs[0], outs[1])
    ref = np.full(n, 9, dtype=dt)
    assert olib().oracle_stencil_raw(C.byref(_stencil_desc(abi.STENCIL_1D3, dt, a0.ctypes.data, ref.ctypes.data, (n,), [1]))) == 0
    assert np.array_equal(ref, outs[0])
    At = torch.from_numpy(a0).cuda()
    out = torch.full((n,), 9, dtype=At.dtype, device="cuda")
    cuda.check(cuda.lib.racu_stencil(C.byref(_stencil_desc(abi.STENCIL_1D3, dt, At.data_ptr(), out.data_ptr(), (n,), [1])), cuda.stream_ptr()))
    assert cuda.lib.racu_last_kernel() == b"stencil_1d3"
    assert np.array_equal(out.cpu().numpy(), ref)
    # a source that does not start on a 16-byte boundary: not for this kernel (the lane-wise map kernel takes it), same answer
    if n > 8:
        big = torch.zeros(n + 1, dtype=At.dtype, device="cuda")
        big[1:] = At
        outu = torch.full((n,), 9, dtype=At.dtype, device="cuda")
        cuda.check(cuda.lib.racu_stencil(C.byref(_stencil_desc(abi.STENCIL_1D3, dt, big[1:].data_ptr(), outu.data_ptr(), (n,), [1])), cuda.stream_ptr()))
        assert cuda.lib.racu_last_kernel() != b"stencil_1d3"
        assert np.array_equal(outu.cpu().numpy(), ref)
    # a destination with a stride: element-wise stores
    out2 = torch.full((2 * n,), 9, dtype=At.dtype, device="cuda")
    s = _stencil_desc(abi.STENCIL_1D3, dt, At.data_ptr(), out2.data_ptr(), (n,), [1])
    s.dst_stride[0] = 2
    cuda.check(cuda.lib.racu_stencil(C.byref(s), cuda.stream_ptr()))
    o2 = out2.cpu().numpy()
    assert np.array_equal(o2[0::2], ref) and bool((o2[1::2] == 9).all())
