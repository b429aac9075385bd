import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_oracle():
    """The C oracle is test infrastructure; build it (and oracle/_ref where the reference exists) once."""
    import harness
    harness.build_oracle()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(HERE, "golden")


def _cuda_device_present():
    if os.environ.get("IPIC_FORCE_GPU_TESTS"):
        return True
    return any(os.path.exists(d) for d in ("/dev/nvidia0", "/dev/nvidiactl", "/dev/dxg"))


def pytest_collection_modifyitems(config, items):
    """GPU tests never run by accident on a CPU-only box: without a CUDA device node they are skipped (with the reason
    shown), whatever -m expression was given.  The product itself still fails loudly without CUDA (test_capi_library.py)."""
    if _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device on this machine (GPU parity tests run on the B200 box with -m gpu)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
