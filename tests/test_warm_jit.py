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
