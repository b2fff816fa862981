import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def fire_lib():
    """Builds (if stale) and loads libfire_b200.so; GPU tests go through it, never through a fallback."""
    from fire_b200 import _lib, build
    build.build()
    return _lib.lib()


@pytest.fixture(scope="session")
def oracle_native():
    from oracle import native
    native.build()
    return native
