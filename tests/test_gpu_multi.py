"""Multi-GPU parity (needs >= 2 GPUs: run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).
Slab-sharded stencil over the library's NCCL communicator must equal the single-GPU run bit for bit; sharded folds
combine with racu_allreduce."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, shape, steps, out_dir, mode="nccl"):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)  # control plane only; data goes over the library's NCCL comm
    from ra_ra_b200 import abi, cuda, parallel, ra
    lib = cuda.lib
    comm = parallel.make_comm(lib, cuda.check, dist, world, rank)
    a0 = np.random.default_rng(41).uniform(-1, 1, shape).astype(np.float32)
    sl = parallel.SlabStencil3D(shape, world, rank)
    if mode == "push":  # one kernel per step: sweep + halo push into the neighbours' ghosts over peer memory
        slab = cuda.PeerSlab(dist, world, rank, sl.local_shape, abi.F32)
        slab.bufs[0].copy_(torch.from_numpy(a0[sl.global_slice()].copy()))
        torch.cuda.synchronize()
        dist.barrier()
        sweep_push = cuda.stencil_sweep_push()
        for t in range(steps):
            sl.step_push(t, slab.ptrs[:2], slab.ptrs[2], slab.peers, 4, sweep_push)
        torch.cuda.synchronize()
        assert slab.timed_out() == 0
        assert cuda.lib.racu_last_kernel() == b"stencil_tma<3D7, halo push>"
        dist.barrier()
        bufs = [slab.bufs[steps % 2].clone()]
        slab.close()
    else:
        A = torch.from_numpy(a0[sl.global_slice()].copy()).cuda()
        An = torch.zeros_like(A)
        bufs = [A, An]
        sweep = cuda.stencil_sweep()
        exch = parallel.RacuCommExchanger(lib, comm, abi.F32, cuda.stream_ptr, cuda.check)
        ov = cuda.StreamOverlap()
        for _ in range(steps):
            sl.step(bufs[0].data_ptr(), bufs[1].data_ptr(), 4, sweep, exch, ov)
            bufs.reverse()
        torch.cuda.synchronize()
    np.save(os.path.join(out_dir, f"A_{rank}.npy"), bufs[0][sl.owned_local()].cpu().numpy())
    # sharded fold: local partial + racu_allreduce
    ex = cuda.CudaExecutor(rank)
    x = torch.from_numpy(a0[sl.lo:sl.hi].copy()).cuda()
    d, keep = ra.flatten(ra.sqrm(ex.wrap(x)), ex=ex)
    out = torch.zeros(1, device="cuda", dtype=torch.float32)
    ex.reduce_dev(d, abi.RED_SUM, out)
    cuda.check(lib.racu_allreduce(comm, C.c_void_p(out.data_ptr()), 1, abi.F32, abi.RED_SUM, cuda.stream_ptr()))
    torch.cuda.synchronize()
    np.save(os.path.join(out_dir, f"sqrm_{rank}.npy"), out.cpu().numpy())
    cuda.check(lib.racu_comm_destroy(comm))
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("mode", ["nccl", "push"])
@pytest.mark.parametrize("shape", [(64, 40, 132), (33, 64, 64), (8, 300, 260)])
def test_sharded_stencil_equals_single_gpu(shape, mode, tmp_path):
    import torch
    import torch.multiprocessing as mp
    from ra_ra_b200 import cuda
    world, steps = min(_ngpu(), 4), 7
    port = 29900 + (os.getpid() + shape[0] + (40 if mode == "push" else 0)) % 90
    mp.spawn(_worker, args=(world, port, shape, steps, str(tmp_path), mode), nprocs=world, join=True)
    a0 = np.random.default_rng(41).uniform(-1, 1, shape).astype(np.float32)
    A = torch.from_numpy(a0).cuda()
    An = torch.zeros_like(A)
    bufs = [A, An]
    sweep = cuda.stencil_sweep()
    for _ in range(steps):
        sweep(bufs[0].data_ptr(), bufs[1].data_ptr(), shape, list(A.stride()), 0, 0)
        bufs.reverse()
    ref = bufs[0].cpu().numpy()
    got = np.concatenate([np.load(tmp_path / f"A_{r}.npy") for r in range(world)])
    assert np.array_equal(got, ref)
    s = float(np.load(tmp_path / "sqrm_0.npy")[0])
    r64 = float((a0.astype(np.float64) ** 2).sum())
    assert abs(s - r64) <= 1e-5 * r64
