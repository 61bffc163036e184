"""Run-time specialised kernels (csrc/jit.cc) against the CPU oracle on the same inputs: the same typed postfix program is
executed by the oracle's sequential ply_ravel restatement and by the NVRTC-instantiated kernel bodies. Maps, integer work
and amin/amax folds are bit-exact; float sums are compared within 8 eps sum|x| (the order of a fold is unspecified,
docs/ra-ra.texi:3102)."""
import numpy as np
import pytest

from jit_cases import CASES, M, N, make_inputs
from oracle_exec import OracleExecutor
from ra_ra_b200 import abi, ra

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def exc():
    from ra_ra_b200 import cuda
    return cuda.CudaExecutor()


def _np(ex, v):
    return v.owner.cpu().numpy().copy() if ex.name == "cuda" else np.array(v.owner, copy=True)


@pytest.mark.parametrize("name", sorted(CASES))
def test_jit_kernels_match_oracle(name, exc):
    from ra_ra_b200 import cuda
    kind, make = CASES[name]
    outs = []
    for ex in (OracleExecutor(), exc):
        q = make_inputs(ex)
        e, dst, assign, redop = make(q)
        if kind == "fold":
            d, keep = ra.flatten(e, ex=ex)
            outs.append(np.float64(ex.reduce(d, redop, abi.F32)))
        else:
            root = dst
            dst._assign(assign, e)
            outs.append(_np(ex, root))
        if ex.name == "cuda":
            assert cuda.lib.racu_last_kernel().decode().startswith("jit_"), cuda.lib.racu_last_kernel()
    if kind == "map":
        assert np.array_equal(outs[0], outs[1], equal_nan=True)
    elif name == "fold_amax_where":
        assert outs[0] == outs[1]
    else:  # float sums: |diff| <= 8 eps * (number of terms * largest magnitude involved)
        tol = 8 * np.finfo(np.float32).eps * N * 64.0 * (M if kind == "fold" else 1)
        assert np.all(np.abs(outs[0] - outs[1]) <= tol), np.abs(outs[0] - outs[1]).max()


def test_jit_is_cached_and_matches_the_interpreter(exc, monkeypatch):
    """Second use of a program neither recompiles nor reloads (cheap calls), and the specialised kernel writes the same bits as the
    interpreting kernel on a large map. (Speed is measured in profiles/, not asserted here.)"""
    import time
    import torch
    from ra_ra_b200 import cuda
    n = 1 << 24
    a, b, c = (exc.wrap(torch.rand(n, device="cuda")) for _ in range(3))
    e = ra.where(a < c, a * c, ra.sqrt(c)) + 0.5

    def run():
        b.assign(e)
        torch.cuda.synchronize()
    run()
    assert cuda.lib.racu_last_kernel() == b"jit_map"
    t0 = time.perf_counter()
    for _ in range(5):
        run()
    t_jit = (time.perf_counter() - t0) / 5
    assert t_jit < 0.05, t_jit  # a recompilation takes seconds, a reload milliseconds
    ref = b.owner.clone()
    monkeypatch.setenv("RACU_NO_JIT", "1")
    run()
    assert cuda.lib.racu_last_kernel().decode().startswith("ply_map")
    assert torch.equal(ref, b.owner)  # same bits from both kernels
