import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu():
    """GPU tests are skipped only where the library loads and sees no device.  A missing, stale or
    ABI-mismatched libtisph.so is an error, not a reason to skip: _capi.load() raises and the
    collection fails loudly."""
    import ti_sph_b200
    return ti_sph_b200.load().tisph_device_count() > 0


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
