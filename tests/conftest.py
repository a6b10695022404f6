import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100a) device; run on the B200 box with -m gpu")


@pytest.fixture(scope="session")
def lib():
    """The built CUDA library (loading it needs no GPU)."""
    from arxiv_rag_b200 import _lib

    if not _lib.LIB_PATH.exists():
        from arxiv_rag_b200.build import build

        build()
    return _lib.lib()


@pytest.fixture(scope="session")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
