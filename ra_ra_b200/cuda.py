"""Executor over the CUDA library (ra_ra_b200/libracu.so): the product path.

Importing this module loads the library and fails loudly if it has not been built (python -m ra_ra_b200.build, or
__graft_entry__.build()). There is no CPU fallback: every map / reduction here is a kernel launch through the C ABI.
PyTorch is used only for device memory and streams.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import abi, ra

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libracu.so")
if not os.path.exists(_PATH):
    raise ImportError(f"{_PATH} is missing: build the CUDA library first (python -m ra_ra_b200.build). "
                      "There is no CPU fallback for the ply path.")
lib = abi.bind(C.CDLL(_PATH))

_TORCH = {abi.BOOL: torch.bool, abi.I32: torch.int32, abi.I64: torch.int64, abi.F32: torch.float32, abi.F64: torch.float64}
_RACU = {v: k for k, v in _TORCH.items()}


def check(st):
    if st != abi.OK:
        raise ra.RaError(st, lib.racu_last_error_string().decode())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class CudaExecutor:
    name = "cuda"

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("ra_ra_b200.cuda needs a CUDA device; there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)

    def wrap(self, t):
        """View over a torch CUDA tensor (any strides)."""
        assert t.is_cuda
        return ra.View(self, t, t.data_ptr(), _RACU[t.dtype], list(zip(t.shape, t.stride())))

    def empty(self, shape, dtype):
        return self.wrap(torch.empty(tuple(shape), dtype=_TORCH[dtype], device=self.device))

    def from_numpy(self, arr):
        return self.wrap(torch.from_numpy(np.array(arr, order="C", copy=True)).to(self.device))  # (ascontiguousarray would promote 0-d to 1-d)

    def to_numpy(self, v):
        out = self.empty(v.shape, v.dtype)
        out.assign(v)
        return out.owner.cpu().numpy()

    def map(self, d):
        check(lib.racu_map(C.byref(d), stream_ptr()))

    def agree(self, d):
        rank = C.c_int32()
        lens = (C.c_int64 * abi.MAX_RANK)()
        check(lib.racu_agree(C.byref(d), C.byref(rank), lens))
        return [lens[k] for k in range(rank.value)]

    def reduce(self, d, redop, rtype):
        out = np.zeros(2, dtype=ra.RACU2NP[rtype])
        check(lib.racu_reduce(C.byref(d), redop, out.ctypes.data, stream_ptr()))
        return out[0]

    def reduce_dev(self, d, redop, out_tensor):
        check(lib.racu_reduce_dev(C.byref(d), redop, C.c_void_p(out_tensor.data_ptr()), stream_ptr()))

    def reduce_into(self, d, redop, ptr):
        """Whole-expression reduction into device memory (no host round trip): racu_reduce_dev."""
        check(lib.racu_reduce_dev(C.byref(d), redop, C.c_void_p(ptr), stream_ptr()))

    def last_kernel(self):
        return lib.racu_last_kernel().decode()


class StreamOverlap:
    """fork()/back()/join() for parallel.SlabStencil3D.step: the halo exchange runs on a side stream while the planes that
    read no ghost are swept on the main stream."""

    def __init__(self):
        self.main = torch.cuda.current_stream()
        self.side = torch.cuda.Stream()

    def fork(self):
        self.main = torch.cuda.current_stream()
        self.side.wait_stream(self.main)
        torch.cuda.set_stream(self.side)

    def back(self):
        torch.cuda.set_stream(self.main)

    def join(self):
        self.main.wait_stream(self.side)


def stencil_sweep(kind=abi.STENCIL_3D7, dtype=abi.F32):
    """sweep(src_ptr, dst_ptr, shape, strides, lo0, hi0) through racu_stencil, for parallel.SlabStencil3D."""
    def sweep(src_ptr, dst_ptr, shape, strides, lo0, hi0):
        s = abi.StencilDesc()
        s.kind, s.dtype, s.rank = kind, dtype, len(shape)
        s.src, s.dst = src_ptr, dst_ptr
        for k in range(len(shape)):
            s.len[k], s.src_stride[k], s.dst_stride[k] = shape[k], strides[k], strides[k]
        s.lo0, s.hi0 = lo0, hi0
        check(lib.racu_stencil(C.byref(s), stream_ptr()))
    return sweep


def stencil_steps(a, anext, steps, kind=abi.STENCIL_3D7, scratch=None):
    """`steps` ping-pong sweeps a -> anext -> a ... (bench/bench-stencil3.cc:55-64 with its std::swap, :117-121) on two torch tensors of the
    same shape through racu_stencil_steps: two steps per pass where the TMA kernel applies. Returns the tensor that holds the last state
    (anext when steps is odd); the other one holds the state before, like the reference's loop leaves them."""
    assert a.shape == anext.shape and a.dtype == anext.dtype and a.stride() == anext.stride()
    s = abi.StencilDesc()
    s.kind, s.dtype, s.rank = kind, (abi.F32 if a.element_size() == 4 else abi.F64), a.dim()
    s.src, s.dst = a.data_ptr(), anext.data_ptr()
    for k in range(a.dim()):
        s.len[k], s.src_stride[k], s.dst_stride[k] = a.shape[k], a.stride(k), anext.stride(k)
    check(lib.racu_stencil_steps(C.byref(s), int(steps), scratch.data_ptr() if scratch is not None else None, stream_ptr()))
    return anext if steps & 1 else a


def stencil_sweep_push(kind=abi.STENCIL_3D7, dtype=abi.F32):
    """sweep_push(...) through racu_stencil_push, for parallel.SlabStencil3D.step_push."""
    def sweep_push(src_ptr, dst_ptr, shape, strides, push_lo, push_hi, flags, signal_lo, signal_hi, step):
        s = abi.StencilDesc()
        s.kind, s.dtype, s.rank = kind, dtype, len(shape)
        s.src, s.dst = src_ptr, dst_ptr
        for k in range(len(shape)):
            s.len[k], s.src_stride[k], s.dst_stride[k] = shape[k], strides[k], strides[k]
        h = abi.HaloPush()
        h.push_lo, h.push_hi, h.flags, h.signal_lo, h.signal_hi, h.step = push_lo, push_hi, flags, signal_lo, signal_hi, step
        check(lib.racu_stencil_push(C.byref(s), C.byref(h), stream_ptr()))
    return sweep_push


class PushStepper:
    """Steps of parallel.SlabStencil3D.step_push with everything but the step number prepared once: the two (descriptor,
    halo) pairs for even / odd steps. One call = one racu_stencil_push launch."""

    def __init__(self, sl, slab, kind=abi.STENCIL_3D7, dtype=abi.F32):
        self.t, self.pairs = 0, []
        es = abi.DTYPE_SIZE[dtype]
        for par in (0, 1):
            got = []
            sl.step_push(par, slab.ptrs[:2], slab.ptrs[2], slab.peers, es, lambda *a: got.append(a))
            src, dst, shape, strides, push_lo, push_hi, flags, sig_lo, sig_hi, _ = got[0]
            s, h = abi.StencilDesc(), abi.HaloPush()
            s.kind, s.dtype, s.rank, s.src, s.dst = kind, dtype, 3, src, dst
            for k in range(3):
                s.len[k], s.src_stride[k], s.dst_stride[k] = shape[k], strides[k], strides[k]
            h.push_lo, h.push_hi, h.flags, h.signal_lo, h.signal_hi = push_lo, push_hi, flags, sig_lo, sig_hi
            self.pairs.append((s, h, C.byref(s), C.byref(h)))

    def step(self):
        s, h, ps, ph = self.pairs[self.t & 1]
        h.step = self.t
        check(lib.racu_stencil_push(ps, ph, stream_ptr()))
        self.t += 1


class _RawCuda:
    """Zero-copy torch view of a racu_alloc'ed buffer (through __cuda_array_interface__)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}


class PeerSlab:
    """The two slab buffers (A / Anext, ghosts included) and the flag block of one rank, allocated with racu_alloc so that
    they can be exported with CUDA IPC, plus the mappings of the two neighbours' (racu_ipc_open). `dist` is only the control
    plane that carries the 64-byte handles."""

    def __init__(self, dist, world, rank, local_shape, dtype=abi.F32):
        self.dist, self.world, self.rank = dist, world, rank
        es = abi.DTYPE_SIZE[dtype]
        nbytes = int(np.prod(local_shape)) * es
        self.ptrs = []
        for nb in (nbytes, nbytes, abi.HALO_FLAG_WORDS * 4):
            p = C.c_void_p()
            check(lib.racu_alloc(C.byref(p), nb))
            self.ptrs.append(p.value)
        typestr = "<f4" if dtype == abi.F32 else "<f8"
        self.bufs = [torch.as_tensor(_RawCuda(p, local_shape, typestr), device="cuda") for p in self.ptrs[:2]]
        self.flags = torch.as_tensor(_RawCuda(self.ptrs[2], (abi.HALO_FLAG_WORDS,), "<i4"), device="cuda")
        self.flags.zero_()
        for b in self.bufs:
            b.zero_()
        torch.cuda.synchronize()
        mine = []
        for p in self.ptrs:
            h = (C.c_ubyte * abi.IPC_HANDLE_BYTES)()
            check(lib.racu_ipc_export(C.c_void_p(p), h))
            mine.append(bytes(h))
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        self.peers, self._opened = {}, []
        for r in (rank - 1, rank + 1):
            if 0 <= r < world:
                mapped = []
                for h in everyone[r]:
                    q = C.c_void_p()
                    check(lib.racu_ipc_open(h, C.byref(q)))
                    mapped.append(q.value)
                    self._opened.append(q.value)
                self.peers[r] = tuple(mapped)
        dist.barrier()

    def timed_out(self):
        return int(self.flags[4].item())

    def close(self):
        torch.cuda.synchronize()
        for q in self._opened:
            check(lib.racu_ipc_close(C.c_void_p(q)))
        self._opened = []
        self.bufs, self.flags = [], None
        self.dist.barrier()  # nobody frees a buffer a neighbour still maps
        for p in self.ptrs:
            check(lib.racu_free(C.c_void_p(p)))
        self.ptrs = []
