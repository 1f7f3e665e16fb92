import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:  # noqa: BLE001
        return 0


@pytest.fixture(scope="session")
def gpu_count():
    return _gpu_count()


def pytest_collection_modifyitems(config, items):
    # a gpu-marked test on a box without a GPU is an error of invocation, not something to silently pass
    pass
