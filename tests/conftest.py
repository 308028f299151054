import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU restatement (test infrastructure).  Built on first use."""
    from oracle import pyoracle
    pyoracle.lib()
    pyoracle.set_libm_mode(1)
    return pyoracle


@pytest.fixture(scope="session")
def ctx():
    """Device context of the product library; only GPU tests may request it."""
    import moosecad_b200 as mc
    return mc.Context(0)
