import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The CUDA library is a build product (git-ignored): compile it for sm_100a if this checkout
    has not been built yet (nvcc cross-compiles without a GPU; __graft_entry__.build() does the same)."""
    from daskr_b200 import build as b

    srcs = [os.path.join(b.CSRC, f) for f in os.listdir(b.CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    srcs.append(os.path.join(ROOT, "include", "daskr_b200.h"))
    if not os.path.exists(b.LIB) or any(os.path.getmtime(s) > os.path.getmtime(b.LIB) for s in srcs):
        b.build_cuda()
    yield


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (oracle/), built on demand.  Test infrastructure only."""
    from tests import oracle_py

    return oracle_py.load()


@pytest.fixture(scope="session")
def dk():
    """The product: daskr_b200 bound to cuda:0.  Fails loudly if the CUDA library is missing."""
    import daskr_b200

    daskr_b200.lib()
    daskr_b200.init(0)
    yield daskr_b200


@pytest.fixture(autouse=True)
def _default_options(request):
    """Library options are process-wide (like the reference's COMMON blocks): every GPU test starts from the
    defaults, whatever the test before it left behind -- e.g. solve_web(fused=0) leaves the callback form of datv
    selected, whose ||v1|| is summed in another order (rounding-level, not bitwise, agreement with the fused form)."""
    if "dk" in request.fixturenames:
        d = request.getfixturevalue("dk")
        d.set_option(d.OPT_FUSED, 2)
        d.set_option(d.OPT_DEVICE_IO, 0)
        d.set_option(d.OPT_MAX_STEPS, 0)
        d.set_option(d.OPT_ASYNC_OUT, 0)
        d.set_option(99, 0)
    yield
