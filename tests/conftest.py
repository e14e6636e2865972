import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def cpu():
    """Our C restatement of the reference (oracle/cpu_ref.c)."""
    from oracle import oracle
    return oracle.cpu()


@pytest.fixture(scope="session")
def ref():
    """The reference's own AVX2/SSE4 loops (oracle/_ref); skip if never built."""
    from oracle import oracle
    lib = oracle.ref()
    if lib is None:
        pytest.skip("oracle/_ref/libarrowgo_ref.so not built (needs /root/reference)")
    return lib


@pytest.fixture(scope="session")
def isa():
    from oracle import oracle
    return oracle.host_isa()


@pytest.fixture(scope="session")
def ag():
    """The product library through its C ABI.  GPU tests fail loudly if it cannot run."""
    from arrow_go_b200 import _native as N
    N.call("ag_init", -1)
    return N
