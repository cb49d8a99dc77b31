import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def b200():
    """The product library bound to cuda:0 (fails loudly when there is no sm_100 device)."""
    import svt_av1_psy_b200 as pkg
    pkg.init(0)
    yield pkg.dsp
    pkg.shutdown()


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    return o


@pytest.fixture(scope="session")
def refc(oracle, request):
    """ctypes handle on the unmodified reference objects.  On a GPU run (-m gpu) a missing oracle is an ERROR -- the
    parity tests must not silently skip there; the CPU suite may run on a clone without /root/reference."""
    if oracle.ref is None:
        if "gpu" in (request.config.getoption("-m") or ""):
            pytest.fail("oracle/_ref/libsvtav1_ref.so is missing: build it with `python __graft_entry__.py --oracle` where "
                        "/root/reference exists (it travels to the GPU box with the repository)")
        pytest.skip("oracle/_ref/libsvtav1_ref.so not built (no /root/reference)")
    return oracle.ref
