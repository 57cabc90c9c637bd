import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def reference_losses():
    """The reference's own losses.py, loaded by file path; skip where it is absent."""
    from tests._refload import load_reference_losses
    mod = load_reference_losses()
    if mod is None:
        pytest.skip("reference checkout not present on this machine")
    return mod
