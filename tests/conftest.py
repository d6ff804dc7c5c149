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
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def ob():
    import onsas_jl_b200
    return onsas_jl_b200


@pytest.fixture(scope="session")
def hostsim():
    import subprocess
    d = os.path.join(ROOT, "tests", "hostsim")
    subprocess.check_call(["make", "-C", d], stdout=subprocess.DEVNULL)
    from tests.hostsim import hostsim_py
    return hostsim_py
