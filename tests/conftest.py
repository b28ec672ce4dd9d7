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
    """The CPU oracle (test infrastructure), arithmetic routed through the reference-compiled CPU
    backend when oracle/_ref is available, else through the port."""
    from oracle.oracle import get_oracle
    o = get_oracle()
    if not o.use_ref("O0"):
        o.use_port()
    o.set_mode("exact")
    return o


@pytest.fixture(scope="session")
def cg():
    """The product API over libcg_b200.so (GPU tests only)."""
    import computational_graph_b200 as pkg
    return pkg.get_api()
