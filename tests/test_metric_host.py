"""Host side of the metric plugin surface for the names that have no kernel class of their own (`inner_product`, `true_angular`;
fast_distance_alternatives, distances.py:2176-2184), on the oracle-backed engine (CPU): the product's NNDescent -- name resolution,
tree type, no data normalisation, kernel metric id, host-side correction, output dtype -- reproduces the REFERENCE's graph from the
reference's own leaves (goldens of tests/golden/make_golden_metrics2.py)."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, METRIC_KWDS, load_golden


@pytest.mark.parametrize("name", ["latent_inner_1500x20", "latent_trueang_1500x20", "latent_dot_1500x20", "latent_minkowski3_1500x20"])
def test_product_host_path_reproduces_reference_graph(name, monkeypatch):
    import pynndescent_b200.pynndescent_ as mod
    from oracle_engine import OracleEngine
    from pynndescent_b200 import NNDescent
    monkeypatch.setattr(mod, "Engine", OracleEngine)
    G = load_golden(name); metric = GOLDEN_CASES[name]
    idx = NNDescent(G["X"], metric, metric_kwds=METRIC_KWDS.get(name), n_neighbors=int(G["k"]), random_state=int(G["seed"]),
                    n_jobs=int(G["T"]), leaf_array=G["leaf_array"])
    gi, gd = idx.neighbor_graph
    assert idx._angular_trees is (metric == "dot")
    assert idx._engine.metric == {"inner_product": "inner_product", "true_angular": "cosine", "dot": "dot", "minkowski": "minkowski"}[metric]
    assert np.array_equal(gi, G["graph_idx"])
    assert gd.dtype == G["graph_dist"].dtype == (np.float32 if metric == "minkowski" else np.float64)
    np.testing.assert_allclose(gd, G["graph_dist"], rtol=1e-5, atol=1e-6)
    assert idx.c_hist == [int(c) for c in G["c_hist"]]


def test_metric_kwds_surface(monkeypatch):
    """metric_kwds.values() are positional arguments of the distance function in the reference (pynndescent_.py:1064-1065): one
    parameter for minkowski (any key name, like the reference), default p = 2; anything a GPU metric cannot take raises."""
    import pynndescent_b200.pynndescent_ as mod
    from oracle_engine import OracleEngine
    from pynndescent_b200 import NNDescent, distances
    monkeypatch.setattr(mod, "Engine", OracleEngine)
    assert distances.metric_param("minkowski", None) == 2.0 and distances.metric_param("minkowski", {"whatever": 1.5}) == 1.5
    assert distances.metric_param("euclidean", None) is None and distances.metric_param("euclidean", {}) is None
    with pytest.raises(NotImplementedError):
        distances.metric_param("euclidean", {"p": 3})
    with pytest.raises(NotImplementedError):
        distances.metric_param("minkowski", {"p": 3, "w": 1})
    with pytest.raises(ValueError):
        distances.metric_param("minkowski", {"p": -1.0})
    G = load_golden("latent_minkowski3_1500x20")
    with pytest.raises(NotImplementedError):
        NNDescent(G["X"][:200], "manhattan", metric_kwds={"p": 3.0}, n_neighbors=5, random_state=1)
    # minkowski with the default p = 2 orders neighbours like euclidean (same graph indices)
    a = NNDescent(G["X"][:400], "minkowski", n_neighbors=6, random_state=3, n_jobs=8).neighbor_graph
    b = NNDescent(G["X"][:400], "euclidean", n_neighbors=6, random_state=3, n_jobs=8).neighbor_graph
    assert np.array_equal(a[0], b[0])
    np.testing.assert_allclose(a[1], b[1], rtol=1e-6, atol=1e-7)
