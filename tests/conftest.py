import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


@pytest.fixture(scope="session", autouse=True)
def _build_oracle_port():
    # the plain-C oracle is test infrastructure: (re)build it on demand
    from oracle import portapi
    portapi.build()


@pytest.fixture(scope="session", autouse=True)
def _build_product_if_missing():
    """libpdgpu.so / libpdhost.so are build artefacts (git-ignored): build them when absent so that
    a fresh checkout can run the suite. An existing library is never rebuilt here."""
    import subprocess
    lib = os.path.join(ROOT, "pd_mg_pin_corrosion_b200", "libpdgpu.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "pd_mg_pin_corrosion_b200", "csrc"), "-j8"],
                              stdout=subprocess.DEVNULL)
    if not os.path.exists(os.path.join(ROOT, "host", "libpdhost.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "host")], stdout=subprocess.DEVNULL)
