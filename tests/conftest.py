import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


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


@pytest.fixture()
def cg():
    """The product API over libcg_b200.so (GPU tests only), reset to the reference-faithful defaults."""
    import computational_graph_b200 as pkg
    api = pkg.get_api()
    api.set_mode("exact")
    api.set_tf32(False)
    api.set_fix_act_grad(False)
    api.set_fix_broadcast_sum(False)
    api.set_fix_sub_sign(False)
    api.set_fix_tied_params(False)
    api.set_fuse(True)
    api.set_cuda_graph(True)
    api.set_strict_gemm_max(64)
    api.set_ew_jit_min(32768)
    api.set_io_graph(True)
    api.set_tf32_pair(True)
    api.set_hoist(True)
    api.set_vm_max_dim(48)
    api.set_tape_jit(True)
    api.set_elide_root(True)
    return api
