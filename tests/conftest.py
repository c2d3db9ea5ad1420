import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ctx():
    """One bls_ctx on cuda:0.  No skip-on-missing: a GPU test without the CUDA library must fail."""
    import pairing_b200._native as nat
    c = nat.Context(0)
    yield c
    c.close()
