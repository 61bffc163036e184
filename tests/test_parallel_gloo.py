"""Host logic of the multi-GPU path (ra_ra_b200/parallel.py) on CPU: world_size 2 and 3 over gloo, with the oracle as the
local sweep. The slab-sharded stencil must reproduce the single-process result bit for bit, ghosts and boundaries included."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from ra_ra_b200 import abi, parallel  # noqa: E402


def _oracle_sweep(dt):
    from oracle_exec import lib as olib

    def sweep(src_ptr, dst_ptr, shape, strides, lo0, hi0):
        s = abi.StencilDesc()
        s.kind, s.dtype, s.rank = abi.STENCIL_3D7, abi.F32 if dt == np.float32 else abi.F64, 3
        s.src, s.dst = src_ptr, dst_ptr
        for k in range(3):
            s.len[k], s.src_stride[k], s.dst_stride[k] = shape[k], strides[k], strides[k]
        s.lo0, s.hi0 = lo0, hi0
        assert olib().oracle_stencil_raw(C.byref(s)) == 0
    return sweep


class _NoOverlap:
    def fork(self): pass
    def back(self): pass
    def join(self): pass


def _reference(a0, steps, dt):
    sweep = _oracle_sweep(dt)
    A, An = a0.copy(), np.zeros_like(a0)
    st = [s // A.itemsize for s in A.strides]
    for _ in range(steps):
        sweep(A.ctypes.data, An.ctypes.data, A.shape, st, 0, 0)
        A, An = An, A
    return A, An


def _worker(rank, world, port, shape, steps, overlap, dtname, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dt = np.dtype(dtname).type
    rng = np.random.default_rng(41)
    a0 = rng.uniform(-1, 1, shape).astype(dt)
    sl = parallel.SlabStencil3D(shape, world, rank)
    A = a0[sl.global_slice()].copy()
    An = np.zeros_like(A)
    tdt = torch.float32 if dt == np.float32 else torch.float64
    ct = C.c_float if dt == np.float32 else C.c_double

    def view(ptr, count):
        return torch.frombuffer((ct * count).from_address(ptr), dtype=tdt)
    ex = parallel.TorchDistExchanger(dist, rank, world, view)
    sweep = _oracle_sweep(dt)
    for _ in range(steps):
        sl.step(A.ctypes.data, An.ctypes.data, A.itemsize, sweep, ex, _NoOverlap() if overlap else None)
        A, An = An, A
    # also the fold path: per-rank partial + allreduce
    part = torch.tensor([float(a0[sl.lo:sl.hi].astype(np.float64).sum())], dtype=torch.float64)
    dist.all_reduce(part)
    np.save(os.path.join(out_dir, f"A_{rank}.npy"), A[sl.owned_local()])
    np.save(os.path.join(out_dir, f"An_{rank}.npy"), An[sl.owned_local()])
    np.save(os.path.join(out_dir, f"sum_{rank}.npy"), part.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,overlap", [(2, (12, 7, 9), False), (2, (12, 7, 9), True), (3, (11, 6, 8), True), (2, (5, 6, 7), True), (3, (7, 5, 6), True)])
def test_slab_stencil_matches_single_process(world, shape, overlap, tmp_path):
    steps, dt = 4, np.float32
    port = 29600 + (os.getpid() + world * 7 + shape[0]) % 300
    mp.spawn(_worker, args=(world, port, shape, steps, overlap, np.dtype(dt).name, str(tmp_path)), nprocs=world, join=True)
    a0 = np.random.default_rng(41).uniform(-1, 1, shape).astype(dt)
    A, An = _reference(a0, steps, dt)
    gotA = np.concatenate([np.load(tmp_path / f"A_{r}.npy") for r in range(world)])
    gotAn = np.concatenate([np.load(tmp_path / f"An_{r}.npy") for r in range(world)])
    assert np.array_equal(gotA, A) and np.array_equal(gotAn, An)
    s = np.load(tmp_path / "sum_0.npy")[0]
    assert abs(s - a0.astype(np.float64).sum()) < 1e-9


def test_shard_bounds_cover_axis():
    for n in (1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            b = [parallel.shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


@pytest.mark.parametrize("world,shape", [(2, (12, 7, 9)), (3, (11, 6, 8)), (4, (9, 5, 6)), (3, (5, 6, 7)), (8, (16, 5, 5))])
def test_push_schedule_matches_single_process(world, shape):
    """parallel.SlabStencil3D.step_push (the one-kernel-per-step schedule of racu_stencil_push): every pointer it derives
    - neighbour ghost planes in the right buffer, flag words - is exercised on host memory with all ranks in this process,
    stepped one after the other, through the host-memory stand-in of the ABI (oracle/libracu_oracle.so, tests only)."""
    shim = C.CDLL(os.path.join(os.path.dirname(HERE), "oracle", "libracu_oracle.so"))
    shim.racu_stencil_push.argtypes = [C.POINTER(abi.StencilDesc), C.POINTER(abi.HaloPush), C.c_void_p]
    steps, dt = 5, np.float32
    a0 = np.random.default_rng(43).uniform(-1, 1, shape).astype(dt)
    slabs = [parallel.SlabStencil3D(shape, world, r) for r in range(world)]
    bufs = [[a0[sl.global_slice()].copy(), np.zeros(sl.local_shape, dt)] for sl in slabs]
    flags = [np.zeros(abi.HALO_FLAG_WORDS, np.uint32) for _ in range(world)]
    ptrs = [(b[0].ctypes.data, b[1].ctypes.data, f.ctypes.data) for b, f in zip(bufs, flags)]

    def sweep_push(src, dst, lshape, strides, push_lo, push_hi, fl, sig_lo, sig_hi, step):
        s = abi.StencilDesc()
        s.kind, s.dtype, s.rank, s.src, s.dst = abi.STENCIL_3D7, abi.F32, 3, src, dst
        for k in range(3):
            s.len[k], s.src_stride[k], s.dst_stride[k] = lshape[k], strides[k], strides[k]
        h = abi.HaloPush()
        h.push_lo, h.push_hi, h.flags, h.signal_lo, h.signal_hi, h.step = push_lo, push_hi, fl, sig_lo, sig_hi, step
        assert shim.racu_stencil_push(C.byref(s), C.byref(h), None) == 0

    for t in range(steps):
        for r, sl in enumerate(slabs):
            # the lock-step condition the kernel waits for: both neighbours have completed t steps on the shared edge
            assert (not sl.g_lo or flags[r][0] >= t) and (not sl.g_hi or flags[r][1] >= t)
            peers = {q: ptrs[q] for q in (r - 1, r + 1) if 0 <= q < world}
            sl.step_push(t, ptrs[r][:2], ptrs[r][2], peers, 4, sweep_push)
    A, An = _reference(a0, steps, dt)
    cur, oth = steps % 2, (steps + 1) % 2
    gotA = np.concatenate([bufs[r][cur][sl.owned_local()] for r, sl in enumerate(slabs)])
    gotAn = np.concatenate([bufs[r][oth][sl.owned_local()] for r, sl in enumerate(slabs)])
    assert np.array_equal(gotA, A) and np.array_equal(gotAn, An)
    for r, sl in enumerate(slabs):
        assert flags[r][0] == (steps if sl.g_lo else 0) and flags[r][1] == (steps if sl.g_hi else 0)


def test_shard0_map_and_fold_equal_the_unsharded_result():
    """B += A*C + v and sum-rows / sum-cols / dot evaluated slab by slab (what each rank of an N-GPU job does) equal the whole-array
    result; the per-slab partials of the folds add up (the allreduce)."""
    sys.path.insert(0, HERE)
    from oracle_exec import OracleExecutor
    from ra_ra_b200 import ra
    ex = OracleExecutor()
    rng = np.random.default_rng(5)
    m, n = 13, 8
    a, b, c = (rng.integers(-4, 5, (m, n)).astype(np.float64) for _ in range(3))
    v = rng.integers(-4, 5, m).astype(np.float64)
    whole = b + a * c + v[:, None]
    for world in (1, 2, 3, 5):
        B = ex.from_numpy(b)
        A, Cc, V = ex.from_numpy(a), ex.from_numpy(c), ex.from_numpy(v)
        rows, dot = np.zeros(n), 0.
        cols = ex.from_numpy(np.zeros(m))
        for rank in range(world):
            sA, sB, sC, sV = (parallel.shard0(x, world, rank) for x in (A, B, Cc, V))
            sB += sA * sC + sV
            part = ex.from_numpy(np.zeros(n))
            pr = part(ra.insert(1))
            pr += sA                                         # sum-rows: every rank holds a full-length partial -> allreduce
            rows += part.numpy()
            sc = parallel.shard0(cols, world, rank)          # sum-cols: outputs stay sharded
            sc += sA
            dot += float(ra.dot(sA, sC))
        assert np.array_equal(B.numpy(), whole)
        assert np.array_equal(rows, a.sum(0)) and np.array_equal(cols.numpy(), a.sum(1)) and dot == float((a * c).sum())
