"""GPU query-walk parity (seam 3): prepared search structures produced by the REAL reference
(tests/golden/make_golden_query.py) -> nnd_query -> the reference's query() answers."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _index(G):
    from pynndescent_b200.search import SearchIndexGPU
    return SearchIndexGPU(G["data"], str(G["metric"]), G["indptr"], G["indices"], G["hyperplanes"], G["offsets"], G["children"],
                          G["tree_indices"], G["vertex_order"], float(G["min_distance"]), int(G["n_neighbors"]),
                          G["search_rng_state"])


@pytest.mark.parametrize("name", ["query_euclid_3k32", "query_cosine_2k24"])
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
