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
    from oracle_lib import Oracle, build_oracle
    build_oracle()
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """The reference's own KDTree.cpp/RansacFilter.cpp built into oracle/_ref (dev container only)."""
    from oracle_lib import Ref
    if not Ref.available():
        pytest.skip("oracle/_ref/libvbref.so not built (needs /root/reference; run `make -C oracle ref`)")
    return Ref()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "correspondence_cv2_4_13.npz"))


@pytest.fixture(autouse=True)
def _reset_vb_options(request):
    """Options set with ctx.set_option (forced code paths) never leak from one test into the next."""
    yield
    if "ctx" in request.fixturenames:
        try:
            request.getfixturevalue("ctx").reset_options()
        except Exception:
            pass
