"""RACU_WARM_JIT=1 python -m pytest tests -m "not gpu": leaves the specialised kernels the GPU tests will ask for in the run-time specialiser's
disk cache (ra_ra_b200/jit_cache: git-ignored, but it travels to the GPU box with the tree), so that `pytest -m gpu` spends its time on the
device and not in NVRTC (the known-answer tests on a B200: 300 s with a cold cache, 10 s with a warm one). Nothing here needs a GPU: the oracle
half of every seeded random case of tests/test_gpu_parity.py runs on the CPU and each descriptor is also handed to racu_prepare (plan +
specialise, nothing launched). Skipped unless RACU_WARM_JIT is set: a cold cache costs about ten minutes of NVRTC."""
import os

import pytest

from oracle_exec import OracleExecutor

pytestmark = pytest.mark.skipif(not os.environ.get("RACU_WARM_JIT"), reason="set RACU_WARM_JIT=1 to warm the specialiser's disk cache")


@pytest.fixture(autouse=True)
def _prepare(monkeypatch):
    monkeypatch.setenv("RACU_ORACLE_PREPARE", "1")


@pytest.mark.parametrize("seed", range(60))
def test_warm_random_maps(seed):
    import test_gpu_parity as gp
    gp.case_random_maps(seed, (OracleExecutor(),))


@pytest.mark.parametrize("seed", range(40))
def test_warm_random_folds(seed):
    import test_gpu_parity as gp
    gp.case_random_folds(seed, (OracleExecutor(),))


@pytest.mark.parametrize("seed", range(40))
def test_warm_random_scatter(seed):
    import test_gpu_parity as gp
    gp.case_random_scatter(seed, (OracleExecutor(),))


def _param_sets(fn):
    """Cartesian product of the pytest.mark.parametrize marks of a test function, as dicts."""
    import itertools
    marks = [m for m in getattr(fn, "pytestmark", []) if m.name == "parametrize"]
    names = [[n.strip() for n in m.args[0].split(",")] for m in marks]
    for combo in itertools.product(*[m.args[1] for m in marks]):
        kw = {}
        for ns, val in zip(names, combo):
            if len(ns) == 1:
                kw[ns[0]] = val
            else:
                kw.update(dict(zip(ns, val)))
        yield kw


@pytest.mark.parametrize("name", ["test_shifted_slice_expressions_run_specialised", "test_user_stencils_on_odd_rows", "test_mass_stencil_on_the_tma_kernel",
                                  "test_contiguous_map_sizes", "test_static_nd_maps_bitexact"])
def test_warm_expression_tests(name):
    """The expression tests of tests/test_gpu_parity.py that run `for ex in (oracle, cuda)`: here with the oracle in both seats (their kernel-name
    assertions only apply to the cuda executor), every parameter set."""
    import test_gpu_parity as gp
    fn = getattr(gp, name)
    for kw in _param_sets(fn):
        fn(exc=OracleExecutor(), **kw)
