"""Sharding of the traversal along its leading axis across the GPUs of one box: one process per GPU.

The reference is single process (README.md:81). What shards: maps and the non-transposing view copies are
independent per slab (no exchange); folds combine per-GPU partials with one allreduce; stencil sweeps exchange one halo
plane with each slab neighbour per step. This module is the host logic for that - slab arithmetic, ghost bookkeeping,
the boundary-first schedule - written against two small interfaces so that it runs unchanged on the device (C-ABI
kernels + NCCL through racu_comm_*) and in CPU tests (gloo + the oracle as the sweep, tests only):

    sweep(src_ptr, dst_ptr, shape, strides, lo0, hi0)   one 7-point sweep over planes [lo0, hi0) of a local slab
    exchanger.exchange(send_lo, recv_lo, send_hi, recv_hi, count)   neighbour planes, non periodic ends

Two schedules for the stencil: `step` (exchange the halo planes of the source with NCCL send/recv, sweep the interior
meanwhile, edges last) and `step_push` (ONE kernel per step: the sweep stores its edge planes straight into the neighbours'
ghost planes over NVLink peer memory and keeps the neighbours in lock step with flags, racu_stencil_push):

    sweep_push(src_ptr, dst_ptr, shape, strides, push_lo, push_hi, flags, signal_lo, signal_hi, step)
"""


def shard_bounds(n, world, rank):
    """Contiguous slab [lo, hi) of an axis of length n owned by `rank`; the remainder goes to the first ranks."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard0(view, world, rank):
    """The slab of `view` this rank owns when an expression is sharded along its leading axis (maps, non-transposing view
    copies and folds are independent per slab). By prefix matching every operand of rank >= 1 is sliced the same way - a row-broadcast
    vector is split with the rows - and rank-0 operands are replicated. Transposes that move the sharded axis are not sharded
    (replicas only): they would need an all-to-all."""
    from . import ra
    if view.rank == 0:
        return view
    lo, hi = shard_bounds(view.len(0), world, rank)
    return view(ra.iota(hi - lo, lo))


class SlabStencil3D:
    """bench/bench-stencil3.cc:55-64 on a global (n0, n1, n2) array sharded along axis 0.

    Rank r owns global planes [lo, hi) and stores planes [lo - g_lo, hi + g_hi) where g_lo / g_hi is 1 when a neighbour
    exists on that side (a ghost plane) and 0 at the ends of the global array (the global boundary plane is owned and, as in
    the reference, never written: bench/bench-stencil3.cc:125-126). The local interior [1, local_n0 - 1) is then exactly the
    set of owned planes that the global sweep updates.

    Two buffers A / Anext are swapped after every step like the reference's std::swap(A.cp, Anext.cp); each keeps its own
    boundary values, so results match the single-GPU run bit for bit.
    """

    def __init__(self, shape, world, rank):
        self.shape = tuple(shape)
        self.world, self.rank = world, rank
        self.lo, self.hi = shard_bounds(shape[0], world, rank)
        self.g_lo = 1 if rank > 0 else 0
        self.g_hi = 1 if rank + 1 < world else 0
        self.local_n0 = (self.hi - self.lo) + self.g_lo + self.g_hi
        self.local_shape = (self.local_n0, shape[1], shape[2])
        self.plane = shape[1] * shape[2]

    def global_slice(self):
        """Slice of the global array held locally (ghosts included)."""
        return slice(self.lo - self.g_lo, self.hi + self.g_hi)

    def owned_local(self):
        """Local plane indices of the owned planes."""
        return slice(self.g_lo, self.g_lo + (self.hi - self.lo))

    def halo_planes(self):
        """(send_lo, recv_lo, send_hi, recv_hi) local plane indices; None where there is no neighbour."""
        n = self.local_n0
        send_lo = 1 if self.g_lo else None          # first owned plane
        recv_lo = 0 if self.g_lo else None          # low ghost
        send_hi = n - 2 if self.g_hi else None      # last owned plane
        recv_hi = n - 1 if self.g_hi else None      # high ghost
        return send_lo, recv_lo, send_hi, recv_hi

    def step(self, src_ptr, dst_ptr, itemsize, sweep, exchanger, overlap=None):
        """One global sweep src -> dst on this rank.

        The halo planes of src are exchanged first (they were final after the previous step). With `overlap` (an object
        with fork() / join() that moves the exchange to a side stream) the planes that do not read a ghost are swept while
        the exchange is in flight, and the one or two edge planes afterwards: boundary cells last, interior overlapped.
        """
        n0, n1, n2 = self.local_shape
        strides = (n1 * n2, n2, 1)
        s_lo, r_lo, s_hi, r_hi = self.halo_planes()

        def addr(base, plane):
            return None if plane is None else base + plane * self.plane * itemsize

        def do_exchange():
            exchanger.exchange(addr(src_ptr, s_lo), addr(src_ptr, r_lo), addr(src_ptr, s_hi), addr(src_ptr, r_hi), self.plane)

        if self.world == 1:
            sweep(src_ptr, dst_ptr, self.local_shape, strides, 0, 0)
            return
        if overlap is None:
            do_exchange()
            sweep(src_ptr, dst_ptr, self.local_shape, strides, 0, 0)
            return
        overlap.fork()
        do_exchange()
        overlap.back()
        lo_edge = 2 if self.g_lo else 1           # first plane that reads no ghost
        hi_edge = n0 - 2 if self.g_hi else n0 - 1  # one past the last such plane
        if hi_edge > lo_edge:
            sweep(src_ptr, dst_ptr, self.local_shape, strides, lo_edge, hi_edge)
        overlap.join()
        if self.g_lo and n0 > 2:
            sweep(src_ptr, dst_ptr, self.local_shape, strides, 1, min(2, n0 - 1))
        if self.g_hi and n0 - 2 >= lo_edge and (n0 - 2 > 1 or not self.g_lo):
            sweep(src_ptr, dst_ptr, self.local_shape, strides, n0 - 2, n0 - 1)


    def push_targets(self, t, itemsize, peers):
        """Where step t (src = buffer t % 2, dst = buffer (t+1) % 2) pushes its edge planes and raises its flags.

        `peers` maps a neighbour's rank to (buffer0_ptr, buffer1_ptr, flags_ptr) as addressable from this rank. The lower
        neighbour's HIGH ghost is the last plane of its slab, the higher neighbour's LOW ghost is its plane 0; flag word 0 of
        a rank is written by its lower neighbour, word 1 by its higher one (include/racu.h, racu_halo_push).
        Returns (push_lo, push_hi, signal_lo, signal_hi), None where there is no neighbour."""
        d = (t + 1) % 2
        push_lo = push_hi = sig_lo = sig_hi = None
        if self.g_lo:
            lo = SlabStencil3D(self.shape, self.world, self.rank - 1)
            push_lo = peers[self.rank - 1][d] + (lo.local_n0 - 1) * self.plane * itemsize
            sig_lo = peers[self.rank - 1][2] + 4
        if self.g_hi:
            push_hi = peers[self.rank + 1][d]
            sig_hi = peers[self.rank + 1][2]
        return push_lo, push_hi, sig_lo, sig_hi

    def step_push(self, t, bufs, flags_ptr, peers, itemsize, sweep_push):
        """Step t as one fused launch: sweep bufs[t % 2] -> bufs[(t+1) % 2] over the whole local interior, edge planes pushed
        into the neighbours' ghosts of THEIR destination buffer. Needs no exchange before it: the ghosts of the source were
        pushed by the neighbours during step t-1 (for t = 0 they come with the initial data)."""
        n0, n1, n2 = self.local_shape
        push_lo, push_hi, sig_lo, sig_hi = self.push_targets(t, itemsize, peers)
        sweep_push(bufs[t % 2], bufs[(t + 1) % 2], self.local_shape, (n1 * n2, n2, 1), push_lo, push_hi, flags_ptr, sig_lo, sig_hi, t)


class TorchDistExchanger:
    """Halo exchange over torch.distributed point-to-point ops (gloo on CPU in tests, NCCL on GPUs). Buffers are given as
    1-D tensors by `view(ptr, count)`; the product path on GPUs uses RacuCommExchanger instead."""

    def __init__(self, dist, rank, world, view):
        self.dist, self.rank, self.world, self.view = dist, rank, world, view

    def exchange(self, send_lo, recv_lo, send_hi, recv_hi, count):
        ops = []
        d = self.dist
        if send_lo is not None:
            ops.append(d.P2POp(d.isend, self.view(send_lo, count), self.rank - 1))
            ops.append(d.P2POp(d.irecv, self.view(recv_lo, count), self.rank - 1))
        if send_hi is not None:
            ops.append(d.P2POp(d.isend, self.view(send_hi, count), self.rank + 1))
            ops.append(d.P2POp(d.irecv, self.view(recv_hi, count), self.rank + 1))
        if ops:
            for w in d.batch_isend_irecv(ops):
                w.wait()


class RacuCommExchanger:
    """Halo exchange through the library's own NCCL communicator (racu_halo_exchange: grouped ncclSend / ncclRecv)."""

    def __init__(self, lib, comm, dtype, stream_ptr, check):
        self.lib, self.comm, self.dtype, self.stream_ptr, self.check = lib, comm, dtype, stream_ptr, check

    def exchange(self, send_lo, recv_lo, send_hi, recv_hi, count):
        import ctypes as C
        p = lambda a: C.c_void_p(a) if a is not None else C.c_void_p(0)  # noqa: E731
        self.check(self.lib.racu_halo_exchange(self.comm, p(send_lo), p(recv_lo), p(send_hi), p(recv_hi), count, self.dtype,
                                               self.stream_ptr()))


def make_comm(lib, check, dist, world, rank):
    """Create the library's NCCL communicator: rank 0 makes the unique id, torch.distributed carries it to the others."""
    import ctypes as C
    from . import abi
    comm = C.c_void_p()
    ident = [None]
    if rank == 0:
        buf = (C.c_ubyte * abi.UNIQUE_ID_BYTES)()
        check(lib.racu_comm_unique_id(buf))
        ident = [bytes(buf)]
    dist.broadcast_object_list(ident, src=0)
    check(lib.racu_comm_init_rank(C.byref(comm), ident[0], world, rank))
    return comm
