import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PKG_DIR = os.path.join(ROOT, "unnamed-rust-sdr_b200")
for p in (HERE, ROOT, PKG_DIR):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import sdr_b200
        return sdr_b200.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # a gpu-marked test on a box without a GPU is an error the user should see, but when the whole
    # suite is run unfiltered on the CPU container we skip them instead of failing
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def sdr():
    import sdr_b200
    return sdr_b200


@pytest.fixture(scope="session")
def O():
    import oracle_lib
    oracle_lib.lib()
    return oracle_lib
