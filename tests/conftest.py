import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def lib_options():
    """Set tuning / test options of libocn_b200 (ocn_set_option) for one test; defaults are restored afterwards."""
    from ocn_b200 import _lib

    def setter(**kw):
        for k, v in kw.items():
            _lib.set_option(k, v)

    yield setter
    _lib.reset_options()
