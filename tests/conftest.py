import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "reference_vectors.npz"))


@pytest.fixture(scope="session")
def models():
    """(stylegan_sd, calibrated iresnet50_sd) -- seeded fixtures, cached under .fixture_cache/."""
    from oracle import fixtures
    return fixtures.build_models(cache_dir=os.path.join(ROOT, ".fixture_cache"))
