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


# ---- minkowski / the metric_kwds surface (NND_METRIC_MINKOWSKI; p through nnd_metric_set_param)
@pytest.mark.parametrize("p", [0.5, 1.5, 3.0, 4.0])
def test_minkowski_pair_distances(torch_cuda, oracle, p):
    """bit-exact vs the oracle on three widths (padded and unpadded rows), and the reference's own float32 values on its KAT matrix"""
    kw = {"p": p}
    oracle.set_metric_param("minkowski", kw)
    for d in (5, 37, 130):
        rs = np.random.RandomState(d)
        X = rs.standard_normal((64, d)).astype(np.float32)
        X[5] = 0.0; X[6] = X[7]
        eng, Xh = _engine(X, "minkowski", 4, metric_kwds=kw)
        a = rs.randint(0, 64, 400).astype(np.int32); b = rs.randint(0, 64, 400).astype(np.int32)
        want = np.array([oracle.dist("minkowski", Xh[i], Xh[j]) for i, j in zip(a, b)], np.float32)
        assert np.array_equal(eng.pair_distances(a, b), want)
    g = load_golden("kat_metrics2")
    A = g["real_A"]
    ii, jj = np.meshgrid(np.arange(24), np.arange(24), indexing="ij")
    eng, _ = _engine(A, "minkowski", 4, metric_kwds=kw)
    got = eng.pair_distances(ii.ravel(), jj.ravel()).reshape(24, 24)
    np.testing.assert_allclose(got, g["dist_minkowski_p%g" % p], rtol=2e-7, atol=0)


@pytest.mark.parametrize("p,n,d,k,seed", [(3.0, 2000, 20, 10, 15), (1.5, 1500, 130, 8, 16), (2.0, 1500, 16, 10, 17)])
def test_minkowski_full_build_and_query_bit_exact_vs_oracle(torch_cuda, oracle, p, n, d, k, seed):
    from conftest import latent
    from pynndescent_b200 import NNDescent
    X = latent(n, d, seed)
    kw = None if p == 2.0 else {"p": p}                                      # p = 2 is the reference's default (distances.py:138)
    idx = NNDescent(X, "minkowski", metric_kwds=kw, n_neighbors=k, random_state=seed)
    gi, gd = idx.neighbor_graph
    o = oracle.OracleNNDescent(X, "minkowski", n_neighbors=k, random_state=seed, n_threads=8, forest="level", metric_kwds=kw)
    oi, od = o.neighbor_graph
    assert idx.c_hist == [int(c) for c in o.c_hist[: o.n_iters_run]]
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    Q = latent(64, d, seed + 100)
    qi, qd = idx.query(Q, k=5)
    exact = (np.abs(Q[:, None, :].astype(np.float64) - X[qi].astype(np.float64)) ** p).sum(-1) ** (1.0 / p)
    np.testing.assert_allclose(qd, exact, rtol=1e-5, atol=1e-6)             # the reported distances are the true ones
    assert np.all(np.diff(qd, axis=1) >= 0)
