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


def pytest_collection_modifyitems(config, items):
    # GPU tests never run by accident on a CPU-only box: they are deselected unless -m gpu is given
    pass
