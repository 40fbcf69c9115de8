"""GPU query-walk parity (seam 3): prepared search structures produced by the REAL reference
(tests/golden/make_golden_query.py) -> nnd_query -> the reference's query() answers."""
import numpy as np
import pytest

from conftest import QUERY_GOLDENS, load_golden

pytestmark = pytest.mark.gpu


def _index(G):
    from pynndescent_b200.search import SearchIndexGPU
    return SearchIndexGPU(G["data"], str(G["metric"]), G["indptr"], G["indices"], G["hyperplanes"], G["offsets"], G["children"],
                          G["tree_indices"], G["vertex_order"], float(G["min_distance"]), int(G["n_neighbors"]),
                          G["search_rng_state"])


@pytest.mark.parametrize("name", list(QUERY_GOLDENS[:2]))   # the later goldens: tests/test_xpending_gpu.py
def test_query_matches_reference(name, oracle):
    G = load_golden(name)
    si = _index(G)
    qi, qd = si.query(G["Q"], k=int(G["k_query"]), epsilon=float(G["epsilon"]))
    assert np.array_equal(qi, G["out_idx"])
    np.testing.assert_allclose(qd, G["out_dist"], rtol=1e-5, atol=1e-6)
    assert qd.dtype == G["out_dist"].dtype
    qi2, qd2 = si.query(G["Q"], k=5, epsilon=0.3)
    assert np.array_equal(qi2, G["out_idx_k5_e03"])
    np.testing.assert_allclose(qd2, G["out_dist_k5_e03"], rtol=1e-5, atol=1e-6)
    # bit-exact against the oracle's restatement of search_closure (same canonical arithmetic)
    hi, hd = oracle.search(G["data"], str(G["metric"]), G["indptr"], G["indices"], G["hyperplanes"], G["offsets"], G["children"],
                           G["tree_indices"], G["Q"], int(G["k_query"]), float(G["epsilon"]), float(G["min_distance"]),
                           int(G["n_neighbors"]), G["search_rng_state"])
    si_, sd_ = oracle.deheap_sort(hi, hd)
    assert np.array_equal(qi, G["vertex_order"][si_])
    assert np.array_equal(qd.astype(np.float64), np.asarray(oracle.correction(str(G["metric"]), sd_), np.float64))
    assert si.last_stats["distance_evals"] > 0


def test_query_rng_fallback_and_edge_cases(oracle):
    """Queries that need the shared rng (a leaf smaller than min(k, n_neighbors)) are replayed in query order:
    k larger than the leaf forces that path for every query; also a zero query under cosine stays unfilled."""
    G = load_golden("query_cosine_2k24")
    si = _index(G)
    Q = G["Q"][:40].copy()
    Q[3] = 0.0
    k = int(G["n_neighbors"])          # min(k, n_neighbors) = 10 ... use k=40 with big leaves instead
    qi, qd = si.query(Q, k=40, epsilon=0.2)
    hi, hd = oracle.search(G["data"], "cosine", G["indptr"], G["indices"], G["hyperplanes"], G["offsets"], G["children"],
                           G["tree_indices"], Q, 40, 0.2, float(G["min_distance"]), int(G["n_neighbors"]), G["search_rng_state"])
    si_, sd_ = oracle.deheap_sort(hi, hd)
    want = np.where(si_ >= 0, G["vertex_order"][np.maximum(si_, 0)], -1)
    got = qi.copy()
    assert np.array_equal(got[np.arange(40) != 3], want[np.arange(40) != 3])
    assert np.all(hi[3] == -1)


@pytest.mark.parametrize("name", list(QUERY_GOLDENS[:2]))   # the later goldens: tests/test_xpending_gpu.py
def test_prepare_matches_reference_structures(name):
    """prepare(): from the reference's kNN graph, the GPU hub tree + diversification + the scipy bookkeeping give
    the reference's search graph, search tree, leaf order and min_distance."""
    from pynndescent_b200.engine import Engine
    from pynndescent_b200.prepare import prepare_search_structures
    from pynndescent_b200 import distances as pdist
    import torch
    G = load_golden(name)
    metric = str(G["metric"])
    metric_id, _, _, angular = pdist.resolve(metric)
    eng = Engine(G["X"], metric_id, int(G["n_neighbors"]))
    gi = torch.from_numpy(G["ng_idx"]).to(eng.device); gd = torch.from_numpy(G["ng_dist"]).to(eng.device)
    S = prepare_search_structures(eng, metric, gi, gd, int(G["n_neighbors"]), angular)
    for key in ("vertex_order", "tree_indices", "children", "indptr", "indices"):
        assert np.array_equal(S[key], G[key]), key
    np.testing.assert_allclose(S["hyperplanes"], G["hyperplanes"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(S["offsets"], G["offsets"], rtol=1e-5, atol=1e-5)
    assert np.float32(S["min_distance"]) == np.float32(G["min_distance"])


def test_build_prepare_query_end_to_end(oracle):
    """NNDescent(...).query(): GPU build -> GPU prepare -> GPU walk, against the oracle's prepare + search on the same
    graph, and against brute force (the reference's hub-tree query bar is recall >= 0.90, test_hub_trees.py:197-238)."""
    from pynndescent_b200 import NNDescent
    from conftest import latent
    both = latent(6200, 24, 3)
    X, Q = both[:6000], both[6000:]
    idx = NNDescent(X, "euclidean", n_neighbors=15, random_state=5)
    qi, qd = idx.query(Q, k=10, epsilon=0.2)
    gi, _ = idx.neighbor_graph
    raw_d = idx._engine.heap_dist.cpu().numpy()                      # sorted raw (squared) distances of the built graph
    P = oracle.prepare(X, "euclidean", gi, raw_d, 15, idx.rng_state)
    hi, hd = oracle.search(P["data"], "euclidean", P["indptr"], P["indices"], P["hyperplanes"], P["offsets"], P["children"],
                           P["tree_indices"], Q, 10, 0.2, float(P["min_distance"]), 15, idx.search_rng_state)
    si, sd = oracle.deheap_sort(hi, hd)
    assert np.array_equal(qi, P["vertex_order"][si])
    assert np.array_equal(qd, np.sqrt(sd))
    true = np.argsort(((Q[:, None, :].astype(np.float64) - X[None, :, :]) ** 2).sum(-1), 1)[:, :10]
    rec = np.mean([len(np.intersect1d(true[i], qi[i])) / 10 for i in range(len(Q))])
    assert rec >= 0.90, rec


@pytest.mark.parametrize("metric,n,d,k", [("manhattan", 4000, 20, 12), ("jaccard", 3000, 40, 10), ("hellinger", 3000, 24, 10)])
def test_more_metrics_build_prepare_query(oracle, metric, n, d, k):
    """SURVEY 8f-2 metrics through prepare() and query(): GPU structures and answers == the oracle's on the same graph."""
    from pynndescent_b200 import NNDescent
    from conftest import metric_data
    both = metric_data(metric, n + 150, d, 3)
    X, Q = both[:n], both[n:]
    idx = NNDescent(X, metric, n_neighbors=k, random_state=5)
    qi, qd = idx.query(Q, k=8, epsilon=0.15)
    gi, _ = idx.neighbor_graph
    raw_d = idx._engine.heap_dist.cpu().numpy()
    P = oracle.prepare(X, metric, gi, raw_d, k, idx.rng_state)
    S = idx._search_structures
    assert np.array_equal(S["indptr"], P["indptr"]) and np.array_equal(S["indices"], P["indices"])
    assert np.array_equal(S["vertex_order"], P["vertex_order"]) and np.array_equal(S["children"], P["children"])
    hi, hd = oracle.search(P["data"], metric, P["indptr"], P["indices"], P["hyperplanes"], P["offsets"], P["children"],
                           P["tree_indices"], Q, 8, 0.15, float(P["min_distance"]), k, idx.search_rng_state)
    si, sd = oracle.deheap_sort(hi, hd)
    assert np.array_equal(qi, P["vertex_order"][si])            # -1 (unfilled) wraps like the reference's fancy index
    assert np.array_equal(qd, oracle.correction(metric, sd), equal_nan=True)
