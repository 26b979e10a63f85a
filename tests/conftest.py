import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def ref():
    """The reference's own code (oracle/_ref); skipped where it was never built."""
    from oracle import refjxl
    if not refjxl.available():
        pytest.skip("oracle/_ref not built (needs /root/reference once: make -C oracle ref)")
    return refjxl
