"""pytest configuration.

`-m "not gpu"` : oracle vs golden vectors / reference build, host logic, C-ABI symbol checks (CPU only).
`-m gpu`       : parity tests proper, through the C-ABI on cuda:0, checked against the oracle.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugins-bad_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The checker: the reference's own C when oracle/_ref is built, else our port."""
    import oracle
    return oracle.best()


@pytest.fixture(scope="session")
def port():
    import oracle
    return oracle.get("port")


@pytest.fixture(scope="session")
def vf():
    import b200vf
    return b200vf


@pytest.fixture(scope="session")
def ctx(vf):
    c = vf.Context(0)          # raises loudly without an sm_100 device: no CPU fallback
    yield c
    c.close()


@pytest.fixture()
def rng():
    return np.random.default_rng(0)
