import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def hcs_lib():
    """Builds (if stale) and loads the CUDA library; CPU-only runs may dlopen it but never compute."""
    from mujoco_contact_surfaces_b200 import build
    build.build()
    from mujoco_contact_surfaces_b200 import engine
    return engine.load_library()
