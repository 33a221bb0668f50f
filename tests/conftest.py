import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    from oracle.oracle import Port
    return Port()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    d = os.path.join(ROOT, "tests", "golden")
    return {n: np.load(os.path.join(d, n + ".npz"), allow_pickle=False) for n in ("dense_vector", "halfcircle", "rand2k")}


@pytest.fixture(scope="session")
def have_ref():
    from oracle import oracle
    return oracle.have_ref()


@pytest.fixture(scope="session")
def golden_refgraph():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "refgraph20k.npz"), allow_pickle=False)
