import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Everything native is built in-tree by __graft_entry__.build() (a no-op when up to date)."""
    import __graft_entry__
    __graft_entry__.build()


@pytest.fixture(scope="session")
def port(_built):
    from oracle import pyoracle
    return pyoracle.Port()


@pytest.fixture(scope="session")
def ref(_built):
    """The reference's own Msg.cpp (oracle/_ref/libohref.so); only where it was built or travelled."""
    from oracle import pyoracle
    if not pyoracle.Ref.available():
        pytest.skip("oracle/_ref/libohref.so not present (needs /root/reference at build time)")
    return pyoracle.Ref()


@pytest.fixture(scope="session")
def ctx(_built):
    from ohpipeline_b200 import capi
    c = capi.Context(0)  # raises loudly without a B200: there is no fallback
    yield c
    c.close()
