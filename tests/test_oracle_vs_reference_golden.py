"""Pins the CPU oracle (oracle/nnd_oracle.c) to vectors produced by the REAL reference
(tests/golden/make_golden.py, numba T=8).  CPU-only; runs in the `-m "not gpu"` suite."""
import numpy as np
import pytest

from conftest import FOREST_FROM_REFERENCE, GOLDEN_CASES, METRIC_KATS, METRIC_KWDS, QUERY_GOLDENS, load_golden


def test_tau_rand_kat(oracle):
    g = load_golden("kat")
    for n in "abc":
        s = g[f"tau_{n}_state0"].copy()
        ints = np.array([oracle.tau_rand_int(s) for _ in range(64)], np.int32)
        assert np.array_equal(ints, g[f"tau_{n}_ints"])
        assert np.array_equal(s, g[f"tau_{n}_state1"])
        s = g[f"tau_{n}_state0"].copy()
        fl = np.array([oracle.tau_rand(s) for _ in range(64)], np.float32)
        assert np.array_equal(fl, g[f"tau_{n}_floats"])
    # SURVEY 8a literal values
    s = np.array([-5, 123456789, -987654321], np.int64)
    assert [oracle.tau_rand_int(s) for _ in range(5)] == [-1530850136, 1419355108, -424356903, 1332770144, 572193240]
    assert s.tolist() == [-100400832, 3507296892, -4142669084]


@pytest.mark.parametrize("name,k", [("h5", 5), ("h15", 15), ("h30", 30)])
def test_heap_push_kat(oracle, name, k):
    g = load_golden("kat")
    L = oracle.lib()
    pri = np.full(k, np.inf, np.float32); ind = np.full(k, -1, np.int32); flg = np.zeros(k, np.uint8)
    d, q, f = g[f"{name}_d"], g[f"{name}_q"], g[f"{name}_f"]
    ret = [L.orc_checked_flagged_heap_push(pri, ind, flg, k, float(d[i]), int(q[i]), int(f[i])) for i in range(len(d))]
    assert np.array_equal(np.array(ret, np.int32), g[f"{name}_ret"])
    assert np.array_equal(pri, g[f"{name}_pri"]) and np.array_equal(ind, g[f"{name}_ind"])
    assert np.array_equal(flg, g[f"{name}_flg"])
    si, sd = oracle.deheap_sort(ind.copy()[None, :], pri.copy()[None, :])
    assert np.array_equal(si[0], g[f"{name}_sorted_ind"]) and np.array_equal(sd[0], g[f"{name}_sorted_pri"])
    pri2 = np.full(k, np.inf, np.float32); ind2 = np.full(k, -1, np.int32)
    ret2 = [L.orc_simple_heap_push(pri2, ind2, k, float(d[i]), int(q[i])) for i in range(len(d))]
    assert np.array_equal(np.array(ret2, np.int32), g[f"{name}_simple_ret"])
    assert np.array_equal(pri2, g[f"{name}_simple_pri"]) and np.array_equal(ind2, g[f"{name}_simple_ind"])


def test_heap_push_survey_literal(oracle):
    L = oracle.lib()
    pri = np.full(5, np.inf, np.float32); ind = np.full(5, -1, np.int32); flg = np.zeros(5, np.uint8)
    seq = [(0.5, 7), (0.3, 2), (0.9, 4), (0.3, 9), (0.1, 7), (0.2, 1), (0.05, 3), (0.3, 2), (0.25, 8), (0.9, 11)]
    ret = [L.orc_checked_flagged_heap_push(pri, ind, flg, 5, d, n, 1) for d, n in seq]
    assert ret == [1, 1, 1, 1, 0, 1, 1, 0, 1, 0]
    assert ind.tolist() == [2, 8, 9, 1, 3]
    si, _ = oracle.deheap_sort(ind[None, :].copy(), pri[None, :].copy())
    assert si[0].tolist() == [3, 1, 8, 9, 2]


@pytest.mark.parametrize("metric", ["sqeuclidean", "cosine", "dot", "correlation"])
def test_distance_kat(oracle, metric):
    """Distances agree with the reference within 1e-5 relative (the reference's own low bits
    depend on LLVM's fastmath vectorisation; the oracle has ONE defined order)."""
    g = load_golden("kat")
    A = g["dist_A"]
    D = np.array([[oracle.dist(metric, A[i], A[j]) for j in range(24)] for i in range(24)], np.float32)
    R = g[f"dist_{metric}"]
    big = R > 1e30  # FLOAT32_MAX sentinel (non-positive similarity / zero vector)
    assert np.array_equal(big, D > 1e30)
    np.testing.assert_allclose(D[~big], R[~big], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("metric", list(METRIC_KATS))
def test_distance_kat_more_metrics(oracle, metric):
    """SURVEY 8f-2 metrics vs the reference's own numba functions: the integer-valued ones (hamming, the binary
    family) and max / f64-accumulated sums must be EXACT, the f32-accumulated log2 ones within 1e-5."""
    g = load_golden("kat_metrics")
    A = g[METRIC_KATS[metric]]
    D = np.array([[oracle.dist(metric, A[i], A[j]) for j in range(24)] for i in range(24)], np.float32)
    R = g[f"dist_{metric}"]
    special = ~np.isfinite(R) | (R > 1e30)      # -log2(0) = inf (jaccard), FLOAT32_MAX sentinel (hellinger)
    assert np.array_equal(special, ~np.isfinite(D) | (D > 1e30))
    assert np.array_equal(R[special], D[special])
    if metric in ("jaccard", "hellinger", "canberra"):
        np.testing.assert_allclose(D[~special], R[~special], rtol=1e-5, atol=2e-6)
    else:
        assert np.array_equal(D, R)
    if metric in ("jaccard", "hellinger"):
        fin = np.isfinite(D) & (D < 1e30)
        np.testing.assert_allclose(oracle.correction(metric, D[fin]), g[f"corr_{metric}"][fin], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("metric", ["inner_product", "true_angular"])
def test_distance_kat_alternatives_without_own_class(oracle, metric):
    """inner_product (alternative_inner_product, distances.py:759-790) and true_angular (alternative_cosine + its own correction,
    distances.py:2181-2184) vs the reference's numba functions: sentinel positions identical, values within 1e-5, and the
    corrected distances equal the reference's correction AND its true distance function."""
    g = load_golden("kat_metrics2")
    A = g["real_A"]
    D = np.array([[oracle.dist(metric, A[i], A[j]) for j in range(24)] for i in range(24)], np.float32)
    R = g[f"dist_{metric}"]
    big = R > 1e30
    assert big.any() and np.array_equal(big, D > 1e30) and np.array_equal(D[big], R[big])
    corr = oracle.correction(metric, D)
    if metric == "inner_product":
        # 1 / <x, y> of mixed-sign rows: the f32 sum's error is relative to sum |x_i y_i| (about 25 here), not to a result that
        # cancellation can make 1e-2 -- so the comparison is on the inner product itself, i.e. on the corrected distance
        np.testing.assert_allclose(1.0 / D[~big], 1.0 / R[~big], rtol=1e-5, atol=2e-5)
        np.testing.assert_allclose(corr[~big], g[f"corr_{metric}"][~big], rtol=1e-5, atol=2e-5)
        np.testing.assert_allclose(corr[~big], g[f"true_{metric}"][~big], rtol=1e-5, atol=2e-5)
    else:
        np.testing.assert_allclose(D[~big], R[~big], rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(corr[~big], g[f"corr_{metric}"][~big], rtol=1e-5, atol=1e-6)
        # (the reference's own true_angular(0, 0) is 0.0 while its correction maps alternative_cosine(0, 0) = 0 to 1.0: the graph
        # path reports the correction, so the zero row is left out of the comparison with the true function)
        nz = ~big & (np.abs(A).sum(1)[:, None] > 0) & (np.abs(A).sum(1)[None, :] > 0)
        np.testing.assert_allclose(corr[nz], g[f"true_{metric}"][nz], rtol=2e-4, atol=2e-4)
    if metric == "inner_product":
        assert np.array_equal(corr[big], np.zeros(int(big.sum())))          # the sentinel maps to 0.0 (distances.py:851-852)
        import pynndescent_b200.distances as ours
        assert np.array_equal(ours.correct_alternative_inner_product(D), corr)
    else:
        import pynndescent_b200.distances as ours
        with np.errstate(invalid="ignore"):
            assert np.array_equal(ours.true_angular_from_alt_cosine(D), corr, equal_nan=True)


@pytest.mark.parametrize("p", [0.5, 1.5, 3.0, 4.0])
def test_distance_kat_minkowski(oracle, p):
    """minkowski(x, y, p) (distances.py:138-153) through the oracle's deterministic x ** p = exp2(p * log2 x): the float32 results
    are the reference's own bits on the KAT matrix (zero row, duplicates), for p below and above 1."""
    g = load_golden("kat_metrics2")
    A = g["real_A"]
    oracle.set_metric_param("minkowski", {"p": p})
    D = np.array([[oracle.dist("minkowski", A[i], A[j]) for j in range(24)] for i in range(24)], np.float32)
    R = g["dist_minkowski_p%g" % p]
    np.testing.assert_allclose(D, R, rtol=2e-7, atol=0)
    assert (D == R).mean() > 0.99 and np.all(np.diag(D) == 0)


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_end_to_end_matches_reference(oracle, name):
    """Full build replay: leaf_array, heaps after each init stage, candidates, per-iteration
    accepted-push counts and the final graph are IDENTICAL to the reference's (indices exact,
    distances rtol 1e-5)."""
    G = load_golden(name)
    metric = GOLDEN_CASES[name]
    inject = G["leaf_array"] if name in FOREST_FROM_REFERENCE else None      # see conftest.FOREST_FROM_REFERENCE
    o = oracle.OracleNNDescent(G["X"], metric, n_neighbors=int(G["k"]), random_state=int(G["seed"]), n_threads=int(G["T"]),
                               keep_stages=True, leaf_array=inject, metric_kwds=METRIC_KWDS.get(name))
    assert np.array_equal(o.leaf_array, G["leaf_array"])
    assert np.array_equal(o.heap_after_leaf[0], G["leaf_heap_idx"])
    assert np.array_equal(o.heap_after_leaf[2], G["leaf_heap_flag"])
    assert np.array_equal(o.heap_after_init[0], G["init_heap_idx"])
    assert np.array_equal(o.rng_state_after_init, G["rng_state1"])
    assert np.array_equal(o.c_hist[: o.n_iters_run], G["c_hist"])
    assert np.array_equal(o.heap_final[0], G["final_heap_idx"])
    gi, gd = o.neighbor_graph
    assert np.array_equal(gi, G["graph_idx"])
    np.testing.assert_allclose(gd, G["graph_dist"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", ["cfg1_euclid_1k16", "latent_cosine_2k24"])
def test_stage_candidates_and_join(oracle, name):
    """Stage-injected: reference heap -> oracle candidates == reference candidates (T=8);
    reference candidates -> oracle process_candidates == reference heap after the iteration."""
    G = load_golden(name)
    metric = GOLDEN_CASES[name]
    data = G["X"]
    heap = (G["init_heap_idx"].copy(), G["init_heap_dist"].copy(), G["init_heap_flag"].copy())
    newc, oldc = oracle.new_build_candidates(heap, min(60, int(G["k"])), G["rng_state1"].copy(), int(G["T"]))
    assert np.array_equal(newc, G["it0_new"]) and np.array_equal(oldc, G["it0_old"])
    assert np.array_equal(heap[2], G["it0_flag_after_cand"])
    c, _ = oracle.process_candidates(data, metric, heap, newc, oldc, int(G["T"]))
    assert c == int(G["c_hist"][0])
    assert np.array_equal(heap[0], G["it0_heap_idx"]) and np.array_equal(heap[2], G["it0_heap_flag"])
    # block-size / thread-count invariance of generate+apply (SURVEY hard part 5)
    heap2 = (G["init_heap_idx"].copy(), G["init_heap_dist"].copy(), G["it0_flag_after_cand"].copy())
    c2, _ = oracle.process_candidates(data, metric, heap2, newc, oldc, 3, block_size=10 ** 9)
    assert c2 == c and np.array_equal(heap2[0], heap[0]) and np.array_equal(heap2[1], heap[1])


@pytest.mark.parametrize("name", list(QUERY_GOLDENS))
def test_prepare_matches_reference(oracle, name):
    """_init_search_graph (pynndescent_.py:1333-1662) restated: from the reference's built graph (raw distances) and rng
    state, the oracle's hub tree, diversified + degree-pruned search graph, leaf order and min_distance are the
    reference's own arrays."""
    G = load_golden(name)
    P = oracle.prepare(G["X"], str(G["metric"]), G["ng_idx"], G["ng_dist"], int(G["n_neighbors"]), G["rng_before_prepare"].copy())
    for key in ("indptr", "indices", "children", "tree_indices", "vertex_order"):
        assert np.array_equal(P[key], G[key]), key
    assert np.array_equal(P["data"], G["data"])
    np.testing.assert_allclose(P["hyperplanes"], G["hyperplanes"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(P["offsets"], G["offsets"], rtol=1e-5, atol=1e-6)
    assert np.float32(P["min_distance"]) == np.float32(G["min_distance"])


@pytest.mark.parametrize("name", list(QUERY_GOLDENS))
def test_query_walk_matches_reference(oracle, name):
    """Seam 3: orc_search on the reference's prepared structures reproduces the reference's raw result heaps
    (layout included) and its final query() answers."""
    G = load_golden(name)
    metric = str(G["metric"])
    hi, hd = oracle.search(G["data"], metric, G["indptr"], G["indices"], G["hyperplanes"], G["offsets"], G["children"],
                           G["tree_indices"], G["Q"], int(G["k_query"]), float(G["epsilon"]), float(G["min_distance"]),
                           int(G["n_neighbors"]), G["search_rng_state"])
    assert np.array_equal(hi, G["heap_idx"])
    np.testing.assert_allclose(hd, G["heap_dist"], rtol=1e-5, atol=1e-6)
    si, sd = oracle.deheap_sort(hi, hd)
    assert np.array_equal(G["vertex_order"][si], G["out_idx"])
    np.testing.assert_allclose(oracle.correction(metric, sd), G["out_dist"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("case", ["rand", "knn", "knn_dist"])
@pytest.mark.parametrize("metric", ["euclidean", "cosine"])
def test_warm_start_matches_reference(oracle, case, metric):
    """init_graph / init_dist (pynndescent_.py:1225-1242, utils.py:836-860): the supplied graph replaces leaf init AND
    init_random; init_dist is pushed verbatim.  Final graph identical to the reference's (tests/golden/make_golden_warm.py)."""
    G = load_golden("warm_start_1500x16")
    kw = {"rand": dict(init_graph=G["g_rand"]), "knn": dict(init_graph=G["g_knn"]),
          "knn_dist": dict(init_graph=G["g_knn"], init_dist=G["d_knn"])}[case]
    o = oracle.OracleNNDescent(G["X"], metric, n_neighbors=int(G["k"]), random_state=int(G["seed"]), n_threads=int(G["T"]), **kw)
    gi, gd = o.neighbor_graph
    assert np.array_equal(gi, G[f"{case}_{metric}_idx"])
    np.testing.assert_allclose(gd, G[f"{case}_{metric}_dist"], rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError, match="Init graph size does not match dataset size!"):
        oracle.OracleNNDescent(G["X"], metric, n_neighbors=10, random_state=1, init_graph=G["g_knn"][:-1])
    with pytest.raises(ValueError, match="shapes of init graph and init distances do not match"):
        oracle.OracleNNDescent(G["X"], metric, n_neighbors=10, random_state=1, init_graph=G["g_knn"], init_dist=G["d_knn"][:, :-1])
