"""racu_stencil_steps: T ping-pong steps of the 7-point sweep taken two per pass (the intermediate state only in shared memory) must leave
BOTH buffers exactly as T single steps do (bench/bench-stencil3.cc:55-64 + std::swap :117-121): bit for bit, boundaries included."""
import ctypes as C

import numpy as np
import pytest

from ra_ra_b200 import abi, ra

pytestmark = pytest.mark.gpu


def _desc(dt, src_ptr, dst_ptr, shape, strides, kind=abi.STENCIL_3D7):
    s = abi.StencilDesc()
    s.kind, s.dtype, s.rank = kind, ra.NP2RACU[np.dtype(dt)], len(shape)
    s.src, s.dst = src_ptr, dst_ptr
    for k in range(len(shape)):
        s.len[k] = shape[k]
        s.src_stride[k] = s.dst_stride[k] = strides[k]
    return s


def _oracle_steps(a, b, steps):
    from oracle_exec import lib as olib
    a, b = a.copy(), b.copy()
    st = [a.strides[k] // a.itemsize for k in range(a.ndim)]
    for t in range(steps):
        x, y = (b, a) if t & 1 else (a, b)
        assert olib().oracle_stencil_raw(C.byref(_desc(a.dtype, x.ctypes.data, y.ctypes.data, a.shape, st))) == 0
    return a, b


@pytest.mark.parametrize("shape", [(9, 20, 132), (13, 31, 252), (7, 16, 124), (30, 45, 384), (5, 3, 8), (12, 14, 120), (11, 30, 244)])
@pytest.mark.parametrize("steps", [0, 1, 3, 4, 5, 6, 7, 10])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_steps_equal_the_oracle_on_both_buffers(shape, steps, dt):
    import torch
    from ra_ra_b200 import cuda
    rng = np.random.default_rng(11)
    a = rng.uniform(-1, 1, shape).astype(dt)
    b = rng.uniform(-1, 1, shape).astype(dt)      # different boundary values in the two buffers, as in the reference's benchmark (value / 0)
    ra_, rb_ = _oracle_steps(a, b, steps)
    A, B = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    last = cuda.stencil_steps(A, B, steps)
    assert last is (B if steps & 1 else A)
    assert np.array_equal(A.cpu().numpy(), ra_) and np.array_equal(B.cpu().numpy(), rb_), cuda.lib.racu_last_kernel()


def test_two_step_kernel_is_the_one_that_runs():
    import torch
    from ra_ra_b200 import cuda
    a = torch.rand(20, 40, 256, device="cuda")
    b = torch.zeros_like(a)
    cuda.lib.racu_launch_count_reset()
    cuda.stencil_steps(a, b, 10)
    # 4 passes of two steps + the copy of the faces into the scratch buffer + 2 single steps
    assert cuda.lib.racu_launch_count() == 7, cuda.lib.racu_launch_count()


@pytest.mark.parametrize("dt", ["float32", "float64"])
def test_steps_with_leading_dimensions_and_caller_scratch(dt):
    """views with leading dimensions larger than the extents (still TMA expressible), scratch supplied by the caller"""
    import torch
    from ra_ra_b200 import cuda
    tdt = getattr(torch, dt)
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    big_a = torch.rand(26, 50, 264, device="cuda", generator=g, dtype=tdt)
    big_b = torch.rand(26, 50, 264, device="cuda", generator=g, dtype=tdt)
    scratch = torch.full((26, 50, 264), 9.0, device="cuda", dtype=tdt)
    A, B = big_a[1:25, 2:47, 4:260], big_b[1:25, 2:47, 4:260]
    ra_, rb_ = big_a.clone(), big_b.clone()
    Ar, Br = ra_[1:25, 2:47, 4:260], rb_[1:25, 2:47, 4:260]
    for t in range(9):
        x, y = (Br, Ar) if t & 1 else (Ar, Br)
        cuda.check(cuda.lib.racu_stencil(C.byref(_desc(np.dtype(dt), x.data_ptr(), y.data_ptr(), tuple(A.shape), list(A.stride()))), cuda.stream_ptr()))
    cuda.stencil_steps(A, B, 9, scratch=scratch)
    assert torch.equal(big_a, ra_) and torch.equal(big_b, rb_)     # (everything outside the views untouched as well)


def test_steps_fall_back_to_single_steps():
    """rows that TMA cannot describe (odd row length), a 2-D kind: the same answers through single steps"""
    import torch
    from ra_ra_b200 import cuda
    rng = np.random.default_rng(12)
    a = rng.uniform(-1, 1, (10, 12, 37)).astype(np.float32)
    b = np.zeros_like(a)
    ra_, rb_ = _oracle_steps(a, b, 6)
    A, B = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    cuda.stencil_steps(A, B, 6)
    assert np.array_equal(A.cpu().numpy(), ra_) and np.array_equal(B.cpu().numpy(), rb_)
    a2 = rng.uniform(-1, 1, (40, 64)).astype(np.float64)
    A2, B2 = torch.from_numpy(a2).cuda(), torch.zeros(40, 64, dtype=torch.float64, device="cuda")
    cuda.stencil_steps(A2, B2, 5, kind=abi.STENCIL_2D5)
    x, y = a2.copy(), np.zeros_like(a2)
    for t in range(5):
        s, d = (y, x) if t & 1 else (x, y)
        d[1:-1, 1:-1] = (((-4 * s[1:-1, 1:-1] + s[2:, 1:-1]) + s[1:-1, 2:]) + s[:-2, 1:-1]) + s[1:-1, :-2]
    assert np.array_equal(A2.cpu().numpy(), x) and np.array_equal(B2.cpu().numpy(), y)


def test_steps_reject_overlapping_buffers():
    import torch
    from ra_ra_b200 import cuda
    a = torch.rand(8, 16, 64, device="cuda")
    s = _desc(np.float32, a.data_ptr(), a.data_ptr(), (8, 16, 64), list(a.stride()))
    assert cuda.lib.racu_stencil_steps(C.byref(s), 4, None, cuda.stream_ptr()) == abi.ERR_UNSUPPORTED


def test_steps_reject_a_bad_scratch_buffer_before_touching_the_arrays():
    import torch
    from ra_ra_b200 import cuda
    a = torch.rand(12, 16, 128, device="cuda")
    b = torch.zeros_like(a)
    a0, b0 = a.clone(), b.clone()
    s = _desc(np.float32, a.data_ptr(), b.data_ptr(), (12, 16, 128), list(a.stride()))
    raw = torch.zeros(12 * 16 * 128 + 8, device="cuda")
    assert cuda.lib.racu_stencil_steps(C.byref(s), 7, raw[1:].data_ptr(), cuda.stream_ptr()) == abi.ERR_BAD_DESC      # not 16-byte aligned
    assert cuda.lib.racu_stencil_steps(C.byref(s), 7, b.data_ptr(), cuda.stream_ptr()) == abi.ERR_BAD_DESC            # one of the arrays
    torch.cuda.synchronize()
    assert torch.equal(a, a0) and torch.equal(b, b0)


def test_steps_at_256_cubed_equal_single_steps():
    import torch
    from ra_ra_b200 import cuda
    n = 256
    g = torch.Generator(device="cuda")
    g.manual_seed(41)
    a = torch.rand(n, n, n, device="cuda", generator=g) * 2 - 1
    b = torch.zeros_like(a)
    ra_, rb_ = a.clone(), b.clone()
    for t in range(10):
        x, y = (rb_, ra_) if t & 1 else (ra_, rb_)
        cuda.check(cuda.lib.racu_stencil(C.byref(_desc(np.float32, x.data_ptr(), y.data_ptr(), (n, n, n), list(a.stride()))), cuda.stream_ptr()))
    cuda.stencil_steps(a, b, 10)
    assert cuda.lib.racu_last_kernel() == b"stencil_tma<3D7>"      # the last two steps are single ones
    assert torch.equal(a, ra_) and torch.equal(b, rb_)
