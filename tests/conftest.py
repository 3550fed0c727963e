import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) on a box without a CUDA device.  With a device they always run: a missing
    libjammy_b200.so must fail loudly there (there is no fallback path to hide behind)."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def lib_built():
    """Build libjammy_b200.so when nvcc is available and the library is stale (cross-compiles without a GPU)."""
    from jammy_flows_b200 import build
    if build.needs_build():
        build.build(verbose=False)
    return build.LIB_PATH
