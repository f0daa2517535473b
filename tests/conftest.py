import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as ge
    return ge.load_package()


@pytest.fixture(scope="session")
def vec(pkg):
    return pkg.vectors


@pytest.fixture(scope="session")
def ctx(pkg):
    """One GPU context for the whole session; fails loudly when the CUDA library or the GPU is missing."""
    c = pkg.Context(0)
    yield c
    c.close()


@pytest.fixture(autouse=True)
def _order_torch_and_library_streams(request):
    """GPU tests fill device buffers with torch (legacy default stream) and then hand raw pointers to the library,
    whose own stream is non-blocking: nothing would order the fills before the kernels.  Every GPU test therefore
    starts with the library on torch's current stream (tests that switch streams do so explicitly after this)."""
    if request.node.get_closest_marker("gpu") is not None and "ctx" in request.fixturenames:
        import torch
        c = request.getfixturevalue("ctx")
        torch.cuda.synchronize()
        # torch's default stream has handle 0, which the library reads as "use your own stream": name the legacy
        # default stream explicitly (cudaStreamLegacy = 0x1), the stream torch's fills and copies run on
        c.set_stream(torch.cuda.current_stream().cuda_stream or 1)
    yield


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    d = os.path.join(ROOT, "tests", "golden")
    return {f[:-4]: np.load(os.path.join(d, f)) for f in os.listdir(d) if f.endswith(".npz")}
