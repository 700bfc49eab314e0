import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden():
    import torch
    return torch.load(os.path.join(ROOT, "tests", "golden", "reference_outputs.pt"), map_location="cpu")


@pytest.fixture(scope="session")
def built_library():
    """Build (or reuse) libhicom_b200.so once per session."""
    from hicom_b200 import build
    return build.build()
