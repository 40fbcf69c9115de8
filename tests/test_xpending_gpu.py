"""GPU parity tests written AFTER the last hardware run of round 1 (first run: round-end suite / scripts/r2_measure.sh).  The file
name sorts after every hardware-verified test file, so that `pytest -x -m gpu` reports those first.

Content: `inner_product` (NND_METRIC_ALT_INNER: the dot kernel class with a 1/x tail) and `true_angular` (alternative_cosine kernel,
other host correction) -- pair distances bit-exact vs the oracle, the reference's KAT matrices, the reference's staged builds
(tests/golden/make_golden_metrics2.py) stage by stage, and whole builds bit-exact vs the oracle.  The bodies are the ones of
tests/test_gpu_parity.py, called with the new parameters.
"""
import numpy as np
import pytest

import test_gpu_parity as tp
import test_gpu_query as tq
from conftest import GOLDEN_CASES_PENDING_GPU, QUERY_GOLDENS, load_golden
from test_gpu_parity import _engine, torch_cuda  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("metric", ["inner_product", "true_angular"])
@pytest.mark.parametrize("d", [3, 16, 37, 128, 200])
def test_pair_distances_bit_exact_vs_oracle(torch_cuda, oracle, metric, d):
    tp.test_pair_distances_bit_exact_vs_oracle(torch_cuda, oracle, metric, d)


def test_inner_product_and_true_angular_kat_vs_reference(torch_cuda):
    """alternative_inner_product / alternative_cosine matrices of the reference's numba functions (kat_metrics2.npz): same
    sentinel positions; the corrected distances = the reference's correction of its own matrix (tolerance on the inner product
    itself -- 1 / <x, y> amplifies the f32 sum's rounding when the terms cancel)."""
    from pynndescent_b200 import distances
    g = load_golden("kat_metrics2")
    A = g["real_A"]
    ii, jj = np.meshgrid(np.arange(24), np.arange(24), indexing="ij")
    for metric in ("inner_product", "true_angular"):
        eng, _ = _engine(A, metric, 4)
        got = eng.pair_distances(ii.ravel(), jj.ravel()).reshape(24, 24)
        R = g[f"dist_{metric}"]
        big = R > 1e30
        assert big.any() and np.array_equal(big, got > 1e30) and np.array_equal(got[big], R[big])
        corr = distances.resolve(metric)[2](got)
        np.testing.assert_allclose(corr[~big], g[f"corr_{metric}"][~big], rtol=1e-5, atol=2e-5 if metric == "inner_product" else 1e-6)
        if metric == "inner_product":
            assert np.array_equal(corr[big], np.zeros(int(big.sum())))



@pytest.mark.parametrize("name", list(GOLDEN_CASES_PENDING_GPU))
def test_staged_builds_from_reference_goldens(torch_cuda, oracle, name):
    tp.test_leaf_init_from_reference_leaves(torch_cuda, oracle, name)
    tp.test_init_random_build_candidates_process(torch_cuda, oracle, name)
    tp.test_end_to_end_with_reference_leaves_matches_reference_graph(torch_cuda, name)


@pytest.mark.parametrize("metric,n,d,k,seed", [("inner_product", 2000, 20, 10, 13), ("true_angular", 2000, 24, 10, 14)])
def test_full_build_bit_exact_vs_oracle(torch_cuda, oracle, metric, n, d, k, seed):
    tp.test_more_metrics_full_build_bit_exact_vs_oracle(torch_cuda, oracle, metric, n, d, k, seed)


# query / prepare goldens of manhattan, jaccard and hellinger (tests/golden/make_golden_query.py), added after the last hardware run
@pytest.mark.parametrize("name", list(QUERY_GOLDENS[2:]))
def test_query_and_prepare_match_reference(name, oracle):
    tq.test_query_matches_reference(name, oracle)
    tq.test_prepare_matches_reference_structures(name)
