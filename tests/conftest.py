import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "golden.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def big():
    """Live-reference fixtures of the benchmarked configurations (tests/golden/make_golden_big.py)."""
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "golden_big.npz"), allow_pickle=False)


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The shared library is built in-tree by `__graft_entry__.build()`; build it if it is missing."""
    so = os.path.join(ROOT, "localdiffusion_hallucination_b200", "libld_sampler.so")
    if not os.path.isfile(so):
        import __graft_entry__ as g

        g.build()
    yield
