import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """Build (or reuse) libimm3gpu.so and the oracle.  On the GPU box the prebuilt files are reused."""
    from immutable3_b200 import _build

    try:
        _build.build_lib()
    except RuntimeError:
        if not os.path.exists(_build.LIB):
            raise
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_oracle

    build_oracle.build_oracle()
    yield
