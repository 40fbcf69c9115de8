import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

GOLDEN_CASES = {
    "cfg1_euclid_1k16": "euclidean",
    "latent_euclid_3k32": "euclidean",
    "latent_cosine_2k24": "cosine",
    "latent_corr_1500x20": "correlation",
    "latent_dot_1500x20": "dot",
}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def latent(n, d, seed=0, latent_dim=16):
    """SURVEY.md section 8d 'latent-16' generator (never iid Gaussian: reference recall 0.12)."""
    rs = np.random.RandomState(seed)
    Z = rs.standard_normal((n, latent_dim)).astype(np.float32)
    A = rs.standard_normal((latent_dim, d)).astype(np.float32) / np.sqrt(latent_dim)
    return (Z @ A + 0.05 * rs.standard_normal((n, d)).astype(np.float32)).astype(np.float32)


@pytest.fixture(scope="session")
def oracle():
    from oracle import nnd_oracle
    nnd_oracle.lib()
    return nnd_oracle
