"""The reference's known-answer tests (tests/kat.py) on the CUDA path, through the C ABI."""
import pytest

import kat

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ex():
    from ra_ra_b200 import cuda
    return cuda.CudaExecutor()


@pytest.mark.parametrize("case", kat.CASES, ids=lambda f: f.__name__)
def test_kat_on_cuda(case, ex):
    case(ex)
