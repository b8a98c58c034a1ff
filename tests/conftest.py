import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100) device; run on the B200 box with -m gpu")


@pytest.fixture(scope="session")
def model_root(tmp_path_factory):
    """Directory with the seeded synthetic MANO_RIGHT.pkl / MANO_LEFT.pkl (seed 0)."""
    from ihmr_b200 import synthetic
    d = tmp_path_factory.mktemp("mano_syn")
    synthetic.write_mano_pkls(str(d), seed=0)
    return str(d)


@pytest.fixture(scope="session")
def oracle_layers(model_root):
    from oracle import mano_oracle
    right = mano_oracle.create(os.path.join(model_root, "MANO_RIGHT.pkl"), "mano", use_pca=False, is_rhand=True)
    left = mano_oracle.create(os.path.join(model_root, "MANO_LEFT.pkl"), "mano", use_pca=False, is_rhand=False)
    return right, left
