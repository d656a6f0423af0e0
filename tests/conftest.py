import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def cuda():
    """GPU session set-up shared by every ``-m gpu`` test: the CUDA library must load (no fallback exists) and the
    torch-side oracle runs in strict fp32 (TF32 off) so it is the reference semantics, not a second approximation."""
    import torch

    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device; the product path has no CPU fallback")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from oaprogressionmmf_b200 import _lib

    _lib.load()
    return torch.device("cuda", 0)


@pytest.fixture(autouse=True)
def _no_pipeline_timeouts(request):
    """After every GPU test: no tcgen05/TMA pipeline wait may have timed out (koa_debug_flag)."""
    yield
    if request.node.get_closest_marker("gpu") is None:
        return
    import torch

    from oaprogressionmmf_b200 import _lib

    torch.cuda.synchronize()
    assert _lib.debug_flag() == 0, "a device-side mbarrier wait timed out (pipeline bug)"
