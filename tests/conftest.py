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
    # SURVEY 8f-2 metrics (tests/golden/make_golden_metrics.py): one case per accumulation class
    "latent_manhattan_1500x20": "manhattan",
    "latent_chebyshev_1200x12": "chebyshev",
    "binary_jaccard_1200x40": "jaccard",
    "cat_hamming_1200x30": "hamming",
    "nonneg_hellinger_1200x24": "hellinger",
    # fast_distance_alternatives without a kernel class of their own (tests/golden/make_golden_metrics2.py): inner_product = the dot
    # accumulation with a 1/x tail (mixed-sign data: half the pairs carry the FLOAT32_MAX sentinel), true_angular = alternative_cosine
    "latent_inner_1500x20": "inner_product",
    "latent_trueang_1500x20": "true_angular",
    # metric_kwds surface (pynndescent_.py:1064-1065, 1294-1302): minkowski with p = 3 (METRIC_KWDS below)
    "latent_minkowski3_1500x20": "minkowski",
}
METRIC_KWDS = {"latent_minkowski3_1500x20": {"p": 3.0}}
# goldens whose GPU parity tests have not run on hardware yet: the GPU suite takes them in tests/test_xpending_gpu.py, which sorts
# after every hardware-verified file, so that `pytest -x` reports the verified tests first
GOLDEN_CASES_PENDING_GPU = ("latent_inner_1500x20", "latent_trueang_1500x20", "latent_minkowski3_1500x20")

# Goldens whose RP forest cannot be re-derived bit-exactly by ANY independent implementation: integer / binary
# data under angular trees puts many points exactly on the hyperplane (margin == 0 up to the reference's own
# fastmath rounding noise, rp_trees.py:126-131), so the coin flips -- and with them the whole rng stream -- depend
# on LLVM's summation order.  For these the parity tests start from the reference's leaf_array.
FOREST_FROM_REFERENCE = ("binary_jaccard_1200x40", "cat_hamming_1200x30")

# name -> input matrix key in tests/golden/kat_metrics.npz (pairwise KATs of the reference's numba functions)
METRIC_KATS = {"manhattan": "real_A", "chebyshev": "real_A", "canberra": "real_A", "hamming": "cat_A",
               "jaccard": "bin_A", "dice": "bin_A", "matching": "bin_A", "kulsinski": "bin_A", "rogerstanimoto": "bin_A",
               "sokalmichener": "bin_A", "russellrao": "bin_A", "sokalsneath": "bin_A", "yule": "bin_A",
               "hellinger": "nonneg_A"}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    # a GPU test that hangs (a kernel that never returns) must end the run with a failure, not hold the box: pytest-timeout's
    # thread method terminates the process even when the main thread is blocked inside a CUDA call.  The largest test of the
    # suite (cfg3 at full size: 2M x 256) needs well under a minute of GPU time plus host-side data generation.
    for item in items:
        if item.get_closest_marker("gpu") is not None and item.get_closest_marker("timeout") is None:
            item.add_marker(pytest.mark.timeout(1500, method="thread"))   # > the 900 s the sharded test gives its own subprocess


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def latent(n, d, seed=0, latent_dim=16):
    """SURVEY.md section 8d 'latent-16' generator (never iid Gaussian: reference recall 0.12)."""
    rs = np.random.RandomState(seed)
    Z = rs.standard_normal((n, latent_dim)).astype(np.float32)
    A = rs.standard_normal((latent_dim, d)).astype(np.float32) / np.sqrt(latent_dim)
    return (Z @ A + 0.05 * rs.standard_normal((n, d)).astype(np.float32)).astype(np.float32)


# query goldens (tests/golden/make_golden_query.py): the reference's prepared search structures + its query() answers
QUERY_GOLDENS = ("query_euclid_3k32", "query_cosine_2k24", "query_manhattan_2k20", "query_jaccard_2k40", "query_hellinger_2k24")

# input domains of the SURVEY 8f-2 metrics (same generators as tests/golden/make_golden_metrics.py)
def binary_data(n, d, seed):
    return (latent(n, d, seed) > 0.35).astype(np.float32)


def categorical_data(n, d, seed):
    return np.clip(np.round(latent(n, d, seed) * 1.2), -2, 2).astype(np.float32)


def nonneg_data(n, d, seed):
    X = np.abs(latent(n, d, seed)).astype(np.float32)
    X[X < 0.15] = 0.0
    return X


def metric_data(metric, n, d, seed):
    kind = METRIC_KATS.get(metric, "real_A")
    return {"real_A": latent, "cat_A": categorical_data, "bin_A": binary_data, "nonneg_A": nonneg_data}[kind](n, d, seed)


@pytest.fixture(scope="session")
def oracle():
    from oracle import nnd_oracle
    nnd_oracle.lib()
    return nnd_oracle
