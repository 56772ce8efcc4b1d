import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def ref():
    """The compiled, unmodified reference Library behind extern-C taps (oracle/_ref/libvc2ref.so)."""
    import refapi
    if not refapi.available():
        pytest.skip("oracle/_ref/libvc2ref.so not built (run oracle/build_ref.sh where /root/reference exists)")
    return refapi


@pytest.fixture(scope="session")
def ctx():
    import vc2_reference_b200 as vc2
    c = vc2.Context(0)
    yield c
    c.close()
