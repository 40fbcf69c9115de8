"""Host side of the metric plugin surface for the names that have no kernel class of their own (`inner_product`, `true_angular`;
fast_distance_alternatives, distances.py:2176-2184), on the oracle-backed engine (CPU): the product's NNDescent -- name resolution,
tree type, no data normalisation, kernel metric id, host-side correction, output dtype -- reproduces the REFERENCE's graph from the
reference's own leaves (goldens of tests/golden/make_golden_metrics2.py)."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden


@pytest.mark.parametrize("name", ["latent_inner_1500x20", "latent_trueang_1500x20", "latent_dot_1500x20"])
def test_product_host_path_reproduces_reference_graph(name, monkeypatch):
    import pynndescent_b200.pynndescent_ as mod
    from oracle_engine import OracleEngine
    from pynndescent_b200 import NNDescent
    monkeypatch.setattr(mod, "Engine", OracleEngine)
    G = load_golden(name); metric = GOLDEN_CASES[name]
    idx = NNDescent(G["X"], metric, n_neighbors=int(G["k"]), random_state=int(G["seed"]), n_jobs=int(G["T"]), leaf_array=G["leaf_array"])
    gi, gd = idx.neighbor_graph
    assert idx._angular_trees is (metric == "dot")
    assert idx._engine.metric == {"inner_product": "inner_product", "true_angular": "cosine", "dot": "dot"}[metric]
    assert np.array_equal(gi, G["graph_idx"])
    assert gd.dtype == G["graph_dist"].dtype == np.float64
    np.testing.assert_allclose(gd, G["graph_dist"], rtol=1e-5, atol=1e-6)
    assert idx.c_hist == [int(c) for c in G["c_hist"]]
