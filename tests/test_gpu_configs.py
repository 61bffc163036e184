"""Parity at the FULL sizes of BASELINE.json's configs, through the C ABI on one B200.

What is compared with what, per config:
  cfg1  f32 8192^2  B += A*C + v   bit-exact vs the CPU oracle on 18 sampled 64-row bands (the oracle interprets ~64 M elements too
                                   slowly for a test), and bit-exact over the WHOLE array vs the same ops in the same order and types
                                   evaluated op by op with torch (no contraction possible across separate kernels).
  cfg2  f32 2^30, 32768^2          folds within 8*eps*sum|x| of an fp64 reference (order is unspecified, docs/ra-ra.texi:3102), bit
                                   identical run to run, exact on integer-valued data.
  cfg3  f64 1024^3                 every view assignment element-for-element against index arithmetic (pure movement: exact).
  cfg4  f32 1024^3 x 10 steps      8 sampled planes bit-exact vs the oracle's raw loops run on the slab of planes each one depends on;
                                   an order-independent checksum (int64 sum of the bit patterns) identical run to run.
  cfg5  f64 / f32 16384^2          gemm exact on the integer-valued inputs of bench-gemm.cc:229-230 against the closed form (f64) and
                                   against an exact integer product (f32, values reduced so that fp32 is exact); gemv / gevm exact against
                                   the closed form of bench-gemv.cc:87.
"""
import ctypes as C

import numpy as np
import pytest

from oracle_exec import OracleExecutor
from oracle_exec import lib as olib
from ra_ra_b200 import abi, ra

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def exc():
    from ra_ra_b200 import cuda
    return cuda.CudaExecutor()


def _rand(torch, shape, seed, dtype, lo=-1.0, hi=1.0):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    return torch.rand(shape, generator=g, device="cuda", dtype=dtype) * (hi - lo) + lo


# ---------------------------------------------------------------- cfg1

@pytest.mark.parametrize("vdt", ["f32", "f64"])
def test_cfg1_fused_map_8192_bitexact(vdt, exc):
    """B += A*C + v with v a rank-1 operand matched on the ROWS (prefix rank extension, examples/read-me.cc:32,163)."""
    import torch
    n = 8192
    A_t, B_t, C_t = (_rand(torch, (n, n), 0x5EED0001 + k, torch.float32) for k in range(3))
    v_t = _rand(torch, (n,), 0x5EED0004, torch.float32 if vdt == "f32" else torch.float64)
    B0 = B_t.clone()
    A, B, Cc, v = exc.wrap(A_t), exc.wrap(B_t), exc.wrap(C_t), exc.wrap(v_t)
    B += A * Cc + v
    assert exc.last_kernel() in ("sflat_map", "jit_map")
    # whole array, same ops / order / types, one torch kernel per op: y = y + ((a*c) + v), narrowed to float at the end
    if vdt == "f32":
        ref = B0 + ((A_t * C_t) + v_t[:, None])
    else:
        ref = (B0.double() + ((A_t * C_t).double() + v_t[:, None])).float()
    assert torch.equal(B_t, ref)
    # the oracle on sampled bands
    rows = sorted(set([0, n - 64] + [int(r) for r in np.linspace(0, n - 64, 16).astype(int) // 64 * 64]))
    got = B_t.cpu().numpy()
    a, b0, c, vh = A_t.cpu().numpy(), B0.cpu().numpy(), C_t.cpu().numpy(), v_t.cpu().numpy()
    ox = OracleExecutor()
    for r in rows:
        bb = ox.from_numpy(b0[r:r + 64])
        bb += ox.from_numpy(a[r:r + 64]) * ox.from_numpy(c[r:r + 64]) + ox.from_numpy(vh[r:r + 64])
        assert np.array_equal(bb.owner, got[r:r + 64]), r


def test_cfg1_integer_valued_variant_exact(exc):
    """A = _0 - _1 style operands: every value an exact integer, compared with integer arithmetic."""
    import torch
    n = 8192
    A_t, B_t, C_t = (torch.empty(n, n, device="cuda", dtype=torch.float32) for _ in range(3))
    A, B, Cc = exc.wrap(A_t), exc.wrap(B_t), exc.wrap(C_t)
    A.assign((ra._0 - ra._1) % 64)
    Cc.assign((ra._0 + 2 * ra._1) % 32)
    B.assign(ra._1 % 16)
    B += A * Cc
    i = torch.arange(n, device="cuda", dtype=torch.int64)[:, None]
    j = torch.arange(n, device="cuda", dtype=torch.int64)[None, :]
    c_mod = lambda x, m: torch.fmod(x, m)  # noqa: E731  (C++ % truncates toward zero)
    ref = c_mod(j, 16) + c_mod(i - j, 64) * c_mod(i + 2 * j, 32)
    assert torch.equal(B_t.long(), ref)


# ---------------------------------------------------------------- cfg2

def test_cfg2_reductions_2pow30(exc):
    import torch
    N = 1 << 30
    eps = float(np.finfo(np.float32).eps)
    for lo, hi, seed in ((0.0, 1.0, 0x5EED0011), (-1.0, 1.0, 0x5EED0013)):
        a_t = _rand(torch, (N,), seed, torch.float32, lo, hi)
        b_t = _rand(torch, (N,), seed + 1, torch.float32, lo, hi)
        a, b = exc.wrap(a_t), exc.wrap(b_t)
        chunks = [(k << 26, (k + 1) << 26) for k in range(16)]

        def f64(fn):  # fp64 reference, chunked to bound the temporaries
            return float(sum(fn(a_t[s:e].double(), b_t[s:e].double()).sum() for s, e in chunks))
        sabs = f64(lambda x, y: x.abs())
        for name, fn, ref, bound in (("sum", lambda: ra.sum(a), f64(lambda x, y: x), sabs),
                                     ("sqrm", lambda: ra.reduce_sqrm(a), f64(lambda x, y: x * x), f64(lambda x, y: x * x)),
                                     ("dot", lambda: ra.dot(a, b), f64(lambda x, y: x * y), f64(lambda x, y: (x * y).abs())),
                                     ("sqrm(a-b)", lambda: ra.reduce_sqrm(a - b), f64(lambda x, y: (x.float() - y.float()).double() ** 2),
                                      f64(lambda x, y: (x - y) ** 2))):
            r1 = float(fn())
            assert exc.last_kernel() == "sflat_fold_inner", (name, exc.last_kernel())
            r2 = float(fn())
            assert r1 == r2, name                                   # deterministic: identical bits run to run
            assert abs(r1 - ref) <= 8 * eps * bound, (name, r1, ref, abs(r1 - ref) / bound)
        del a_t, b_t, a, b
        torch.cuda.empty_cache()


def test_cfg2_row_and_column_sums_32768(exc):
    """c[j] = sum_i a[i,j] (bench-sum-rows.cc:26-34) and c[i] = sum_j a[i,j] (bench-sum-cols.cc:23-31), random and integer-valued."""
    import torch
    n = 32768
    eps = float(np.finfo(np.float32).eps)
    a_t = _rand(torch, (n, n), 0x5EED0021, torch.float32)
    A = exc.wrap(a_t)
    for axis, name in ((0, "rows"), (1, "cols")):
        ref = torch.zeros(n, device="cuda", dtype=torch.float64)
        bound = torch.zeros(n, device="cuda", dtype=torch.float64)
        for s in range(0, n, 2048):
            blk = a_t[s:s + 2048].double()
            if axis == 0:
                ref += blk.sum(0)
                bound += blk.abs().sum(0)
            else:
                ref[s:s + 2048] = blk.sum(1)
                bound[s:s + 2048] = blk.abs().sum(1)
        outs = []
        for rep in range(2):
            c_t = torch.zeros(n, device="cuda", dtype=torch.float32)
            c = exc.wrap(c_t)
            if axis == 0:
                cc = c(ra.insert(1))
                cc += A
                assert exc.last_kernel() == "sflat_fold_outer"
            else:
                c += A
                assert exc.last_kernel() == "sflat_fold_inner"
            outs.append(c_t)
        assert torch.equal(outs[0], outs[1]), name
        assert bool(((outs[0].double() - ref).abs() <= 8 * eps * bound).all()), name
    # integer-valued: a[i,j] = ((i - j) % 7) - 3: sums below 2^24, exact in any order
    A.assign((ra._0 - ra._1) % 7 - 3)
    i = torch.arange(n, device="cuda", dtype=torch.int64)
    for axis in (0, 1):
        c_t = torch.zeros(n, device="cuda", dtype=torch.float32)
        c = exc.wrap(c_t)
        if axis == 0:
            cc = c(ra.insert(1))
            cc += A
        else:
            c += A
        ref = torch.zeros(n, device="cuda", dtype=torch.int64)
        for s in range(0, n, 4096):
            blk = torch.fmod(i[s:s + 4096, None] - i[None, :], 7) - 3
            if axis == 0:
                ref += blk.sum(0)
            else:
                ref[s:s + 4096] = blk.sum(1)
        assert torch.equal(c_t.long(), ref), axis


# ---------------------------------------------------------------- cfg3

def _expect_blocks(torch, n, fn, B_t, block=64):
    """B_t[x, y, z] == fn(x, y, z) for every element, in slabs of `block` planes."""
    y = torch.arange(n, device="cuda", dtype=torch.float64)[None, :, None]
    z = torch.arange(n, device="cuda", dtype=torch.float64)[None, None, :]
    for x0 in range(0, n, block):
        x = torch.arange(x0, x0 + block, device="cuda", dtype=torch.float64)[:, None, None]
        if not torch.equal(B_t[x0:x0 + block], fn(x, y, z).expand(block, n, n)):
            return False
    return True


def test_cfg3_strided_views_1024_cubed(exc):
    """transpose / reverse / iota-slice assignment on f64 1024^3 (ra/arrays.hh:408-462, ra/ply.hh:367-435): A[i,j,k] = i*2^20 + j*2^10 + k,
    so any misplaced element is detectable; every assignment is checked element for element."""
    import torch
    n = 1024
    A_t = torch.empty(n, n, n, device="cuda", dtype=torch.float64)
    A = exc.wrap(A_t)
    A.assign(ra._0 * (n * n) + ra._1 * n + ra._2)
    w = (float(n * n), float(n), 1.0)
    assert _expect_blocks(torch, n, lambda x, y, z: x * w[0] + y * w[1] + z * w[2], A_t)
    B_t = torch.empty_like(A_t)
    B = exc.wrap(B_t)
    for perm in ((2, 1, 0), (1, 0, 2), (0, 2, 1), (1, 2, 0)):
        B_t.fill_(-1)
        B.assign(ra.transpose(A, perm))                       # source axis k goes to destination axis perm[k]: B[a] = A[i], i_k = a_perm[k]
        assert _expect_blocks(torch, n, lambda x, y, z: sum((x, y, z)[perm[k]] * w[k] for k in range(3)), B_t), perm
    for k in range(3):
        B_t.fill_(-1)
        B.assign(ra.reverse(A, k))
        assert _expect_blocks(torch, n, lambda x, y, z: sum(((n - 1 - c) if q == k else c) * w[q] for q, c in enumerate((x, y, z))), B_t), k
    # B(iota(n/2, 0, 2), all, iota(n, n-1, -1)) = A(iota(n/2), all, all): even planes reversed along k, odd planes untouched
    B_t.fill_(-1)
    B(ra.iota(n // 2, 0, 2), ra.all, ra.iota(n, n - 1, -1)).assign(A(ra.iota(n // 2), ra.all, ra.all))
    assert _expect_blocks(torch, n, lambda x, y, z: torch.where(x % 2 == 0, (x / 2) * w[0] + y * w[1] + (n - 1 - z) * w[2], torch.full_like(x, -1.0)), B_t)
    # the bench-from pattern with arithmetic index sets (bench/bench-from.cc:84-113, as iota subscripts): a strided box
    B_t.fill_(-1)
    B(ra.iota(n // 4, 1, 4), ra.iota(n // 2, 0, 2), ra.iota(n // 2, 1, 2)).assign(A(ra.iota(n // 4, 3, 2), ra.iota(n // 2, n // 2), ra.iota(n // 2, n - 1, -2)))
    sel = lambda x, y, z: (x % 4 == 1) & (y % 2 == 0) & (z % 2 == 1)  # noqa: E731
    val = lambda x, y, z: (3 + 2 * ((x - 1) / 4)) * w[0] + (n // 2 + y / 2) * w[1] + (n - 1 - 2 * ((z - 1) / 2)) * w[2]  # noqa: E731
    assert _expect_blocks(torch, n, lambda x, y, z: torch.where(sel(x, y, z), val(x, y, z), torch.full_like(x + y + z, -1.0)), B_t)


# ---------------------------------------------------------------- cfg4

def _bits_checksum(torch, t):
    """Order-independent exact checksum: the int64 sum of the 32-bit patterns."""
    return int(t.view(torch.int32).sum(dtype=torch.int64))


def test_cfg4_laplace3d_1024_ten_steps(exc):
    """bench/bench-stencil3.cc:55-64, 10 sweeps with the pointer swap on f32 1024^3. Sampled planes of the result against the oracle's
    raw loops run for 10 steps on the slab of planes (p-11 .. p+11) each sampled plane depends on; the planes at the slab's ends are
    outside the dependence cone of plane p, so what the slab's artificial boundary holds there never reaches it."""
    import torch
    from ra_ra_b200 import cuda
    from test_gpu_parity import _stencil_desc
    n, T, H = 1024, 10, 11
    a0 = _rand(torch, (n, n, n), 0x5EED0041, torch.float32)
    st = list(a0.stride())
    sums = []
    final = None
    for rep in range(2):
        bufs = [a0.clone(), torch.zeros_like(a0)]
        for _ in range(T):
            cuda.check(cuda.lib.racu_stencil(C.byref(_stencil_desc(abi.STENCIL_3D7, np.float32, bufs[0].data_ptr(), bufs[1].data_ptr(), (n, n, n), st)), cuda.stream_ptr()))
            assert cuda.lib.racu_last_kernel() == b"stencil_tma<3D7>"
            bufs.reverse()
        sums.append(_bits_checksum(torch, bufs[0]))
        final = bufs[0]
        del bufs
    assert sums[0] == sums[1]
    for p in (1, 2, 9, 300, 511, 777, n - 3, n - 2):
        lo, hi = max(0, p - H), min(n, p + H + 1)
        cur = a0[lo:hi].cpu().numpy().copy()
        nxt = np.zeros_like(cur)
        shape = cur.shape
        sst = [s // 4 for s in cur.strides]
        for _ in range(T):
            assert olib().oracle_stencil_raw(C.byref(_stencil_desc(abi.STENCIL_3D7, np.float32, cur.ctypes.data, nxt.ctypes.data, shape, sst))) == 0
            cur, nxt = nxt, cur
        assert np.array_equal(final[p].cpu().numpy(), cur[p - lo]), p
    # boundary cells: untouched in both buffers (bench-stencil3.cc:125-126): T even, so `final` is the buffer that started as a0
    assert torch.equal(final[0], a0[0]) and torch.equal(final[-1], a0[-1]) and torch.equal(final[:, 0], a0[:, 0]) and torch.equal(final[:, :, -1], a0[:, :, -1])


# ---------------------------------------------------------------- cfg5

@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_cfg5_gemm_16384_integer_valued_exact(dt, exc):
    """bench/bench-gemm.cc:229-237: a = _0 - _1, b = _1 - 2*_0; every product and partial sum is an integer below 2^53, so f64 is exact in
    any summation order and equals the closed form. For f32 the values are reduced mod 8 so that fp32 (and the 3xTF32 split) is exact too."""
    import torch
    n = 16384
    T = torch.float64 if dt == "f64" else torch.float32
    a_t, b_t, c_t = (torch.empty(n, n, device="cuda", dtype=T) for _ in range(3))
    a, b, c = exc.wrap(a_t), exc.wrap(b_t), exc.wrap(c_t)
    if dt == "f64":
        a.assign(ra._0 - ra._1)
        b.assign(ra._1 - 2 * ra._0)
    else:
        a.assign((ra._0 - ra._1) % 8)
        b.assign((ra._1 - 2 * ra._0) % 8)
    c.assign(1)
    ra.gemm(a, b, c)                                           # c += a b  (ra/ra.hh:403-412)
    assert exc.last_kernel() in ("dgemm_dmma884", "dgemm_tc", "sgemm_3xtf32_tcgen05"), exc.last_kernel()
    i = torch.arange(n, device="cuda", dtype=torch.float64)[:, None]
    if dt == "f64":
        S1, S2 = n * (n - 1) // 2, (n - 1) * n * (2 * n - 1) // 6  # c[i,j] = sum_k (i-k)(j-2k) = n i j - (2i + j) S1 + 2 S2
        for r0 in range(0, n, 2048):
            ii = i[r0:r0 + 2048]
            jj = torch.arange(n, device="cuda", dtype=torch.float64)[None, :]
            ref = 1 + n * ii * jj - (2 * ii + jj) * S1 + 2 * S2
            assert torch.equal(c_t[r0:r0 + 2048], ref), r0
    else:
        for r0 in range(0, n, 4096):                            # exact integer product in fp64 (cuBLAS: the checker, not the product path)
            ref = 1 + a_t[r0:r0 + 4096].double() @ b_t.double()
            assert float(ref.abs().max()) < 2 ** 24
            assert torch.equal(c_t[r0:r0 + 4096].double(), ref), r0


@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_cfg5_gemv_gevm_16384_exact(dt, exc):
    """bench/bench-gemv.cc:87-92: a = _0 - _1, b = 1 - 2*_0, exact (f64: the closed form; f32: values reduced so that fp32 is exact)."""
    import torch
    n = 16384
    T = torch.float64 if dt == "f64" else torch.float32
    a_t = torch.empty(n, n, device="cuda", dtype=T)
    x_t = torch.empty(n, device="cuda", dtype=T)
    a, x = exc.wrap(a_t), exc.wrap(x_t)
    i = torch.arange(n, device="cuda", dtype=torch.float64)
    if dt == "f64":
        a.assign(ra._0 - ra._1)
        x.assign(1 - 2 * ra._0)
        S1, S2 = n * (n - 1) // 2, (n - 1) * n * (2 * n - 1) // 6
        ref_v = n * i - (2 * i + 1) * S1 + 2 * S2            # sum_j (i-j)(1-2j)
        ref_t = S1 - 2 * S2 - i * (n - 2 * S1)               # sum_i (1-2i)(i-j) = S1 - 2 S2 - j (n - 2 S1)
    else:
        a.assign((ra._0 - ra._1) % 8)
        x.assign((1 - 2 * ra._0) % 4)
        ref_v = a_t.double() @ x_t.double()
        ref_t = x_t.double() @ a_t.double()
    y_t = torch.zeros(n, device="cuda", dtype=T)
    ra.gemv(a, x, exc.wrap(y_t))
    assert exc.last_kernel() == "sflat_fold_inner"
    assert torch.equal(y_t.double(), ref_v)
    y_t.zero_()
    ra.gevm(x, a, exc.wrap(y_t))
    assert exc.last_kernel() == "sflat_fold_outer"
    assert torch.equal(y_t.double(), ref_t)


@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_cfg5_gemm_16384_random_tolerance(dt, exc):
    """Random inputs: |c - fp64 reference| <= bound * sum_k |a||b|, bound = 100 eps sqrt(K)-style for f64 (the reference's own BLAS check is
    100 eps, bench-gemm.cc:264) and the 3xTF32 bound (K 2^-24 + 2^-20) for f32; checked on sampled row blocks."""
    import torch
    n = 16384
    T = torch.float64 if dt == "f64" else torch.float32
    a_t = _rand(torch, (n, n), 0x5EED0051, T)
    b_t = _rand(torch, (n, n), 0x5EED0052, T)
    c_t = torch.zeros(n, n, device="cuda", dtype=T)
    ra.gemm(exc.wrap(a_t), exc.wrap(b_t), exc.wrap(c_t))
    bd = b_t.double()
    bound = 100 * np.finfo(np.float64).eps * np.sqrt(n) if dt == "f64" else (n * 2.0 ** -24 + 2.0 ** -20)
    for r0 in (0, 4096 + 64, n - 128):
        ad = a_t[r0:r0 + 128].double()
        ref, mag = ad @ bd, ad.abs() @ bd.abs()
        assert bool(((c_t[r0:r0 + 128].double() - ref).abs() <= bound * mag).all()), r0
