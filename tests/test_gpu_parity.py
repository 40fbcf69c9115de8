"""GPU parity tests (run on the B200 box with `-m gpu`).  Every call goes through the C-ABI
(pynndescent_b200._ffi -> libnnd_b200.so).  Checker = the CPU oracle + golden vectors produced
by the real reference; /root/reference is never read here.

Bars: indices / flags / heap layout / accepted-push counts bit-exact; distances bit-exact against
the oracle (same defined summation order) and within 1e-5 relative of the reference's numbers.
"""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, GOLDEN_CASES_PENDING_GPU, METRIC_KATS, METRIC_KWDS, latent, load_golden, metric_data

# (the cases added after the last hardware run are parametrised in tests/test_xpending_gpu.py; see conftest)
GPU_CASES = [n for n in GOLDEN_CASES if n not in GOLDEN_CASES_PENDING_GPU]

pytestmark = pytest.mark.gpu


class _MetricIds(dict):
    """metric name -> C enum through the product's own registry (pynndescent_b200.distances.resolve)."""
    def __missing__(self, name):
        from pynndescent_b200 import distances
        return distances.resolve(name)[0]


METRIC_ID = _MetricIds()


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _engine(X, metric, k, normalize_dot=True, metric_kwds=None):
    from pynndescent_b200 import distances
    from pynndescent_b200.engine import Engine
    if metric == "dot" and normalize_dot:
        from sklearn.preprocessing import normalize
        X = normalize(X, norm="l2", copy=True)
    return Engine(X, METRIC_ID[metric], k, metric_param=distances.metric_param(metric, metric_kwds)), np.ascontiguousarray(X, np.float32)


def _brute_knn(X, k, rows, metric="euclidean"):
    Xr = X[rows].astype(np.float64)
    if metric == "euclidean":
        d2 = (Xr ** 2).sum(1)[:, None] + (X.astype(np.float64) ** 2).sum(1)[None, :] - 2.0 * Xr @ X.astype(np.float64).T
    else:
        Xn = X / np.maximum(np.linalg.norm(X, axis=1, keepdims=True), 1e-30)
        d2 = 1.0 - (Xn[rows].astype(np.float64) @ Xn.astype(np.float64).T)
    return np.argsort(d2, axis=1, kind="stable")[:, :k]


def _recall(graph_idx, X, k, metric="euclidean", n_sample=500, seed=0):
    rows = np.random.RandomState(seed).choice(X.shape[0], min(n_sample, X.shape[0]), replace=False)
    true = _brute_knn(X, k, rows, metric)
    hit = sum(len(np.intersect1d(true[i], graph_idx[r])) for i, r in enumerate(rows))
    return hit / (len(rows) * k)


# ------------------------------------------------------------------ primitives
def test_rng_stream_matches_reference_kat_and_oracle(torch_cuda, oracle):
    g = load_golden("kat")
    X = np.zeros((4, 4), np.float32)
    eng, _ = _engine(X, "euclidean", 2)
    for n in "abc":
        s0 = g[f"tau_{n}_state0"]
        assert np.array_equal(eng.rng_stream(s0, 0, 64, False), g[f"tau_{n}_ints"])
        assert np.array_equal(eng.rng_stream(s0, 0, 64, True), g[f"tau_{n}_floats"])
        assert np.array_equal(eng.rng_advance(s0, 64), g[f"tau_{n}_state1"])
    # jump-ahead to far offsets == stepping the oracle
    s = np.array([123456789, -987654321, 55555], np.int64)
    ref = s.copy()
    seq = np.array([oracle.tau_rand_int(ref) for _ in range(100000)], np.int32)
    for start, cnt in ((0, 1000), (777, 3000), (65536, 513), (99000, 1000)):
        assert np.array_equal(eng.rng_stream(s, start, cnt, False), seq[start:start + cnt])
    assert np.array_equal(eng.rng_advance(s, 100000), ref)


@pytest.mark.parametrize("metric", ["euclidean", "cosine", "dot", "correlation"])
@pytest.mark.parametrize("d", [3, 16, 37, 128, 200])
def test_pair_distances_bit_exact_vs_oracle(torch_cuda, oracle, metric, d):
    rs = np.random.RandomState(d)
    X = rs.standard_normal((64, d)).astype(np.float32)
    X[5] = 0.0; X[6] = X[7]; X[8] = 1.5
    eng, Xh = _engine(X, metric, 4, normalize_dot=False)
    a = rs.randint(0, 64, 500).astype(np.int32); b = rs.randint(0, 64, 500).astype(np.int32)
    got = eng.pair_distances(a, b)
    want = np.array([oracle.dist(metric, Xh[i], Xh[j]) for i, j in zip(a, b)], np.float32)
    assert np.array_equal(got, want)


def test_distance_kat_vs_reference(torch_cuda):
    g = load_golden("kat")
    A = g["dist_A"]
    ii, jj = np.meshgrid(np.arange(24), np.arange(24), indexing="ij")
    for metric in ("sqeuclidean", "cosine", "dot", "correlation"):
        eng, _ = _engine(A, metric, 4, normalize_dot=False)
        got = eng.pair_distances(ii.ravel(), jj.ravel()).reshape(24, 24)
        R = g[f"dist_{metric}"]
        big = R > 1e30
        assert np.array_equal(big, got > 1e30)
        np.testing.assert_allclose(got[~big], R[~big], rtol=1e-5, atol=2e-6)


# ------------------------------------------------------------------ SURVEY 8f-2 metrics
@pytest.mark.parametrize("metric", list(METRIC_KATS))
def test_more_metrics_pair_distances(torch_cuda, oracle, metric):
    """nnd_pair_distances for every added metric: bit-exact vs the oracle on domain data of three widths, and vs the
    reference's own numba functions on the KAT matrix (exact for the integer-valued / max / f64-summed ones)."""
    for d in (5, 37, 130):
        X = metric_data(metric, 64, d, d)
        X[5] = 0.0; X[6] = X[7]
        eng, Xh = _engine(X, metric, 4)
        rs = np.random.RandomState(d)
        a = rs.randint(0, 64, 400).astype(np.int32); b = rs.randint(0, 64, 400).astype(np.int32)
        got = eng.pair_distances(a, b)
        want = np.array([oracle.dist(metric, Xh[i], Xh[j]) for i, j in zip(a, b)], np.float32)
        assert np.array_equal(got, want, equal_nan=True), metric
    g = load_golden("kat_metrics")
    A = g[METRIC_KATS[metric]]
    ii, jj = np.meshgrid(np.arange(24), np.arange(24), indexing="ij")
    eng, _ = _engine(A, metric, 4)
    got = eng.pair_distances(ii.ravel(), jj.ravel()).reshape(24, 24)
    R = g[f"dist_{metric}"]
    special = ~np.isfinite(R) | (R > 1e30)
    assert np.array_equal(special, ~np.isfinite(got) | (got > 1e30))
    if metric in ("jaccard", "hellinger", "canberra"):
        np.testing.assert_allclose(got[~special], R[~special], rtol=1e-5, atol=2e-6)
    else:
        assert np.array_equal(got, R)


@pytest.mark.parametrize("metric,n,d,k,seed", [("manhattan", 3000, 20, 10, 3), ("chebyshev", 2000, 12, 10, 4),
                                               ("canberra", 2000, 16, 8, 5), ("hamming", 2000, 30, 10, 6),
                                               ("jaccard", 2500, 40, 10, 7), ("dice", 1500, 40, 10, 8),
                                               ("yule", 1500, 33, 6, 9), ("russellrao", 1200, 40, 10, 10),
                                               ("hellinger", 2000, 24, 10, 11), ("kulsinski", 1200, 130, 12, 12)])
def test_more_metrics_full_build_bit_exact_vs_oracle(torch_cuda, oracle, metric, n, d, k, seed):
    """Forest (euclidean or angular per pynndescent_.py:1074-1085) + leaf init + descent under the added metrics ==
    the oracle's level-scheme build, indices and distances bit for bit -- binary / categorical data makes almost every
    comparison a tie, so this also pins the push ORDER."""
    from pynndescent_b200 import NNDescent
    X = metric_data(metric, n, d, seed)
    idx = NNDescent(X, metric, n_neighbors=k, random_state=seed)
    gi, gd = idx.neighbor_graph
    o = oracle.OracleNNDescent(X, metric, n_neighbors=k, random_state=seed, n_threads=8, forest="level")
    oi, od = o.neighbor_graph
    assert idx.c_hist == [int(c) for c in o.c_hist[: o.n_iters_run]]
    assert np.array_equal(gi, oi)
    assert np.array_equal(gd, od, equal_nan=True)


# ------------------------------------------------------------------ stage-injected parity vs the reference
@pytest.mark.parametrize("name", GPU_CASES)
def test_leaf_init_from_reference_leaves(torch_cuda, oracle, name):
    G = load_golden(name); metric = GOLDEN_CASES[name]
    oracle.set_metric_param(metric, METRIC_KWDS.get(name))
    eng, Xh = _engine(G["X"], metric, int(G["k"]), metric_kwds=METRIC_KWDS.get(name))
    eng.leaf_init_from_array(G["leaf_array"])
    hi, hd, hf = eng.get_heap()
    assert np.array_equal(hi, G["leaf_heap_idx"]) and np.array_equal(hf, G["leaf_heap_flag"])
    fin = np.isfinite(G["leaf_heap_dist"])
    np.testing.assert_allclose(hd[fin], G["leaf_heap_dist"][fin], rtol=1e-5, atol=1e-6)
    heap = oracle.make_heap(Xh.shape[0], int(G["k"]))
    oracle.init_rp_tree(Xh, metric, heap, G["leaf_array"])
    assert np.array_equal(hd, heap[1])                       # bit-exact vs the oracle, distances included
    # tiny push capacity -> many routing rounds -> same result (blocking invariance)
    # the routed (counting-sort + ordered apply) leaf path with a tiny push capacity -> many rounds
    import os
    eng.make_heap()
    os.environ["NND_LEAF_ROUTED"] = "1"
    try:
        eng.leaf_init_from_array(G["leaf_array"], push_capacity=1 << 15)
    finally:
        del os.environ["NND_LEAF_ROUTED"]
    hi2, hd2, hf2 = eng.get_heap()
    assert np.array_equal(hi2, hi) and np.array_equal(hd2, hd) and np.array_equal(hf2, hf)


@pytest.mark.parametrize("name", GPU_CASES)
def test_init_random_build_candidates_process(torch_cuda, oracle, name):
    G = load_golden(name); metric = GOLDEN_CASES[name]
    k = int(G["k"]); T = int(G["T"]); C = min(60, k)
    eng, Xh = _engine(G["X"], metric, k, metric_kwds=METRIC_KWDS.get(name))
    eng.set_heap(G["leaf_heap_idx"], G["leaf_heap_dist"], G["leaf_heap_flag"])
    rng = G["rng_state0"].copy()
    eng.init_random(rng)
    assert np.array_equal(rng, G["rng_state1"])
    hi, hd, hf = eng.get_heap()
    assert np.array_equal(hi, G["init_heap_idx"]) and np.array_equal(hf, G["init_heap_flag"])
    for it in range(2):
        if f"it{it}_new" not in G:
            break
        newc, oldc = eng.build_candidates(C, rng, T, tables_valid=(it > 0))
        torch_cuda.cuda.synchronize()
        assert np.array_equal(newc.cpu().numpy(), G[f"it{it}_new"])
        # the OLD lists of vertices without any new candidate are never read (the local join loops over new candidates only,
        # utils.py:575-577) and are left empty here instead of being sampled (k_cand.cu `act`, tests/test_schedule_invariants.py)
        has_new = (G[f"it{it}_new"] >= 0).any(1)
        oc = oldc.cpu().numpy()
        assert np.array_equal(oc[has_new], G[f"it{it}_old"][has_new])
        rest = oc[~has_new]                                # empty (candidate skip) or sampled like the reference (skip off / oracle engine)
        assert ((rest == -1).all(1) | (rest == G[f"it{it}_old"][~has_new]).all(1)).all()
        assert np.array_equal(eng.get_heap()[2], G[f"it{it}_flag_after_cand"])
        c, stats = eng.process_candidates(newc, oldc)
        assert c == int(G["c_hist"][it])
        hi, hd, hf = eng.get_heap()
        assert np.array_equal(hi, G[f"it{it}_heap_idx"]) and np.array_equal(hf, G[f"it{it}_heap_flag"])
        fin = np.isfinite(G[f"it{it}_heap_dist"])
        np.testing.assert_allclose(hd[fin], G[f"it{it}_heap_dist"][fin], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("name", ["cfg1_euclid_1k16", "latent_cosine_2k24"])
def test_row_sharded_candidate_steps_equal_reference_candidates(torch_cuda, name, world):
    """The three steps of the row-sharded candidate build (nnd_cand_local_counts / nnd_cand_export / nnd_cand_import_select,
    csrc/k_cand.cu) for `world` simulated processes on ONE GPU -- every process reads only its own rows' edges, the count tables
    are summed / offset the way sharded.py does it, the 12-byte reverse-edge records are routed by owner -- must reproduce the
    REFERENCE's new_build_candidates output (utils.py:221-320): candidate lists and cleared flags of both golden iterations."""
    G = load_golden(name); metric = GOLDEN_CASES[name]
    k = int(G["k"]); T = int(G["T"]); C = min(60, k)
    eng, Xh = _engine(G["X"], metric, k)
    N = eng.N
    B = N // world + 1
    ranges = [(min(r * B, N), min((r + 1) * B, N)) for r in range(world)]
    rng = G["rng_state1"].copy()
    heaps = [(G["init_heap_idx"], G["init_heap_dist"], G["init_heap_flag"])]
    if "it0_heap_idx" in G and "it1_new" in G:
        heaps.append((G["it0_heap_idx"], G["it0_heap_dist"], G["it0_heap_flag"]))
    for it, (hi_, hd_, hf_) in enumerate(heaps):
        eng.set_heap(hi_, hd_, hf_)
        tabs = []
        for lo, hi in ranges:                                               # step 1 on every "process"
            eng.set_row_range(lo, hi)
            tabs.append(eng.cand_local_counts(C, T, world, B))
        A = np.stack(tabs)
        sent = []                                                           # [src][dst] -> uint8 records
        for r, (lo, hi) in enumerate(ranges):                               # step 2
            eng.set_row_range(lo, hi)
            assert np.array_equal(eng.cand_local_counts(C, T, world, B), tabs[r])       # (shared scratch: recompute this rank's prefix sums)
            glob = np.zeros(6 * T, np.int64)
            for n in range(T):
                glob[6 * n: 6 * n + 4] = A[:, 5 * n: 5 * n + 4].sum(0)
                glob[6 * n + 4] = A[:r, 5 * n + 4].sum(); glob[6 * n + 5] = A[:r, 5 * T].sum()
            cnt = [int(c) for c in A[r, 5 * T + 1:]]
            boff = np.concatenate([[0], np.cumsum(cnt)[:-1]]).astype(np.int64)
            buf = eng.cand_export(C, rng, T, it > 0, world, B, glob, sum(cnt))
            torch_cuda.cuda.synchronize()
            sent.append([buf[int(boff[q]) * 12: (int(boff[q]) + cnt[q]) * 12].clone() for q in range(world)])
        newc = np.full((N, C), -9, np.int32); oldc = np.full((N, C), -9, np.int32)
        for r, (lo, hi) in enumerate(ranges):                               # step 3
            eng.set_row_range(lo, hi)
            eng.cand_local_counts(C, T, world, B)        # (shared scratch: this rank's own new-edge marks for the candidate skip)
            rec = torch_cuda.cat([sent[q][r] for q in range(world)])
            n = rec.numel() // 12
            assert n == int(A[:, 5 * T + 1 + r].sum())
            nc, oc = eng.cand_import_select(C, T, rec if n else torch_cuda.zeros(12, dtype=torch_cuda.uint8, device=eng.device), n)
            torch_cuda.cuda.synchronize()
            newc[lo:hi] = nc.cpu().numpy()[lo:hi]; oldc[lo:hi] = oc.cpu().numpy()[lo:hi]
        eng.set_row_range(0, 2 ** 62)
        has_new = (G[f"it{it}_new"] >= 0).any(1)         # the old lists of vertices without new candidates are left empty (never read)
        assert np.array_equal(newc, G[f"it{it}_new"]) and np.array_equal(oldc[has_new], G[f"it{it}_old"][has_new]) and (oldc[~has_new] == -1).all()
        assert np.array_equal(eng.get_heap()[2], G[f"it{it}_flag_after_cand"])


@pytest.mark.parametrize("name", GPU_CASES)
def test_end_to_end_with_reference_leaves_matches_reference_graph(torch_cuda, name):
    """Same seed, same leaves => the SAME graph as the reference's NNDescent(...) (T=8)."""
    from pynndescent_b200 import NNDescent
    G = load_golden(name); metric = GOLDEN_CASES[name]
    idx = NNDescent(G["X"], metric, metric_kwds=METRIC_KWDS.get(name), n_neighbors=int(G["k"]), random_state=int(G["seed"]),
                    n_jobs=int(G["T"]), leaf_array=G["leaf_array"])
    gi, gd = idx.neighbor_graph
    assert idx.c_hist == [int(c) for c in G["c_hist"]]
    assert np.array_equal(gi, G["graph_idx"])
    np.testing.assert_allclose(gd, G["graph_dist"], rtol=1e-5, atol=1e-6)
    assert gd.dtype == G["graph_dist"].dtype


def test_survey_cfg1_literal_rows(torch_cuda):
    """SURVEY.md 8a golden rows of cfg1 (1k x 16 iid, k=10, random_state=42, T=8)."""
    from pynndescent_b200 import NNDescent
    G = load_golden("cfg1_euclid_1k16")
    idx = NNDescent(G["X"], "euclidean", n_neighbors=10, random_state=42, leaf_array=G["leaf_array"])
    gi, gd = idx.neighbor_graph
    assert gi[0].tolist() == [0, 276, 513, 144, 697, 660, 890, 758, 72, 627]
    assert gi[1].tolist() == [1, 290, 630, 872, 353, 722, 530, 22, 124, 843]
    np.testing.assert_allclose(gd[0, :4], [0, 2.856893, 3.3170645, 3.4720914], rtol=1e-5)


# ------------------------------------------------------------------ forest + full pipeline vs the oracle
@pytest.mark.parametrize("metric,n,d,k", [("euclidean", 5000, 32, 15), ("cosine", 3000, 24, 10),
                                          ("correlation", 2000, 20, 12), ("euclidean", 1000, 5, 10)])
def test_forest_bit_exact_vs_oracle_level_scheme(torch_cuda, oracle, metric, n, d, k):
    X = latent(n, d, 5)
    eng, Xh = _engine(X, metric, k)
    rs = np.random.RandomState(3)
    states = rs.randint(-2 ** 31 + 1, 2 ** 31 - 1, size=(6, 3)).astype(np.int64)
    L = max(60, min(256, 5 * k))
    angular = metric != "euclidean"
    lp, lo, nl, ml = eng.forest(states, L, angular)
    torch_cuda.cuda.synchronize()
    lp, lo = lp.cpu().numpy(), lo.cpu().numpy()
    want = oracle.forest_leaf_array(Xh, states, L, angular, 200, "level")
    assert nl == want.shape[0] and ml == want.shape[1]
    got = np.full_like(want, -1)
    for i in range(nl):
        got[i, : lo[i + 1] - lo[i]] = lp[lo[i]: lo[i + 1]]
    assert np.array_equal(got, want)


@pytest.mark.parametrize("metric,n,d,k,seed", [("euclidean", 6000, 32, 15, 42), ("cosine", 4000, 24, 10, 7),
                                               ("dot", 3000, 20, 8, 5), ("correlation", 3000, 20, 12, 11),
                                               ("euclidean", 2000, 7, 30, 1), ("sqeuclidean", 1500, 130, 5, 2)])
def test_full_build_bit_exact_vs_oracle(torch_cuda, oracle, metric, n, d, k, seed):
    from pynndescent_b200 import NNDescent
    X = latent(n, d, seed)
    if metric == "dot":
        X = np.abs(X)
    idx = NNDescent(X, metric, n_neighbors=k, random_state=seed)
    gi, gd = idx.neighbor_graph
    o = oracle.OracleNNDescent(X, metric, n_neighbors=k, random_state=seed, n_threads=8, forest="level")
    oi, od = o.neighbor_graph
    assert idx.c_hist == [int(c) for c in o.c_hist[: o.n_iters_run]]
    assert np.array_equal(gi, oi)
    assert np.array_equal(gd, od)
    # structural properties the reference's tests pin (test_pynndescent_.py:299-369)
    assert np.all(np.diff(gd, axis=1) >= 0)
    for r in range(0, n, 97):
        assert len(set(gi[r].tolist())) == k
    if metric == "euclidean":
        assert _recall(gi, X, k) >= 0.95


@pytest.mark.parametrize("metric,n,d,k,seed", [("cosine", 30000, 256, 15, 3),        # cfg3 shape: 2 staged d-chunks x 4
                                               ("correlation", 12000, 784, 30, 4),   # cfg5 shape: chunked k_join, C=30
                                               ("euclidean", 20000, 100, 40, 6)])    # C=40 > batched limit, d % 128 != 0
def test_config_shapes_bit_exact_vs_oracle(torch_cuda, oracle, metric, n, d, k, seed):
    """The BASELINE.json config shapes (reduced N) through every kernel variant: multi-chunk rows, the
    chunked one-row-per-CTA join (correlation / large d / large C), routed leaf init for big leaves."""
    from pynndescent_b200 import NNDescent
    X = latent(n, d, seed)
    idx = NNDescent(X, metric, n_neighbors=k, random_state=seed)
    gi, gd = idx.neighbor_graph
    o = oracle.OracleNNDescent(X, metric, n_neighbors=k, random_state=seed, n_threads=8, forest="level")
    oi, od = o.neighbor_graph
    assert idx.c_hist == [int(c) for c in o.c_hist[: o.n_iters_run]]
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)


def test_thread_count_and_capacity_invariance(torch_cuda, oracle):
    """T matters exactly as in the reference (oracle agrees for each T); push capacity never does."""
    from pynndescent_b200 import NNDescent
    X = latent(3000, 16, 9)
    res = {}
    for T in (1, 3, 8):
        a = NNDescent(X, "euclidean", n_neighbors=10, random_state=4, n_jobs=T)
        o = oracle.OracleNNDescent(X, "euclidean", n_neighbors=10, random_state=4, n_threads=T, forest="level")
        assert np.array_equal(a.neighbor_graph[0], o.neighbor_graph[0])
        res[T] = a.neighbor_graph[0]
    b = NNDescent(X, "euclidean", n_neighbors=10, random_state=4, n_jobs=8, push_capacity=1 << 18)
    assert np.array_equal(b.neighbor_graph[0], res[8])


def test_no_tree_init_and_degenerate_data(torch_cuda, oracle):
    from pynndescent_b200 import NNDescent
    X = latent(1500, 12, 2)
    a = NNDescent(X, "euclidean", n_neighbors=10, random_state=3, tree_init=False)
    o = oracle.OracleNNDescent(X, "euclidean", n_neighbors=10, random_state=3, tree_init=False, n_threads=8)
    assert np.array_equal(a.neighbor_graph[0], o.neighbor_graph[0])
    assert np.array_equal(a.neighbor_graph[1], o.neighbor_graph[1])
    # duplicates + zero rows (reference regression data has the same flavour: conftest.py:46-51)
    Y = np.vstack([latent(800, 10, 4), np.zeros((40, 10), np.float32), np.tile(latent(1, 10, 6), (60, 1))]).astype(np.float32)
    for metric in ("euclidean", "cosine"):
        a = NNDescent(Y, metric, n_neighbors=10, random_state=1)
        o = oracle.OracleNNDescent(Y, metric, n_neighbors=10, random_state=1, n_threads=8, forest="level")
        gi, gd = a.neighbor_graph
        assert np.array_equal(gi, o.neighbor_graph[0])
        for r in range(Y.shape[0]):
            row = gi[r][gi[r] >= 0]
            assert len(set(row.tolist())) == len(row)          # no duplicate neighbours in any row


def test_recall_close_to_reference_numbers(torch_cuda):
    """200k x 128 latent-16, k=15: the reference reaches recall 0.911."""
    from pynndescent_b200 import NNDescent
    X = latent(200000, 128, 0)
    idx = NNDescent(X, "euclidean", n_neighbors=15, random_state=42)
    gi, gd = idx.neighbor_graph
    rec = _recall(gi, X, 15, n_sample=1000)
    assert rec >= 0.911 - 0.005, rec
    # size-independent properties at scale
    assert np.all(gi[:, 0] == np.arange(X.shape[0]))            # every point is its own nearest neighbour
    assert np.all(gd[:, 0] == 0) and np.all(np.diff(gd, axis=1) >= 0)
    rows = np.random.RandomState(1).choice(X.shape[0], 2000, replace=False)
    exact = np.sqrt(((X[rows, None, :].astype(np.float64) - X[gi[rows]].astype(np.float64)) ** 2).sum(-1))
    np.testing.assert_allclose(gd[rows], exact, rtol=1e-5, atol=1e-6)
    srt = np.sort(gi[rows], axis=1)
    assert np.all(srt[:, 1:] != srt[:, :-1])


def test_nan_input_raises_like_check_array(torch_cuda):
    """pynndescent_.py:1054 check_array rejects NaN/inf; here the scan runs on the device."""
    from pynndescent_b200 import NNDescent
    X = latent(500, 8, 1)
    X[17, 3] = np.nan
    with pytest.raises(ValueError, match="NaN"):
        NNDescent(X, "euclidean", n_neighbors=5, random_state=0)


@pytest.mark.parametrize("case", ["rand", "knn", "knn_dist"])
@pytest.mark.parametrize("metric", ["euclidean", "cosine"])
def test_warm_start_init_graph(torch_cuda, oracle, case, metric):
    """init_graph / init_dist (pynndescent_.py:1225-1242): nnd_pair_distances + ordered nnd_apply_records build the start
    heap; the final graph equals the reference's (golden) and the oracle's bit for bit; tiny push capacity -> same."""
    from pynndescent_b200 import NNDescent
    G = load_golden("warm_start_1500x16")
    kw = {"rand": dict(init_graph=G["g_rand"]), "knn": dict(init_graph=G["g_knn"]),
          "knn_dist": dict(init_graph=G["g_knn"], init_dist=G["d_knn"])}[case]
    idx = NNDescent(G["X"], metric, n_neighbors=int(G["k"]), random_state=int(G["seed"]), n_jobs=int(G["T"]), **kw)
    gi, gd = idx.neighbor_graph
    assert np.array_equal(gi, G[f"{case}_{metric}_idx"])
    np.testing.assert_allclose(gd, G[f"{case}_{metric}_dist"], rtol=1e-5, atol=1e-6)
    o = oracle.OracleNNDescent(G["X"], metric, n_neighbors=int(G["k"]), random_state=int(G["seed"]), n_threads=int(G["T"]), **kw)
    oi, od = o.neighbor_graph
    assert idx.c_hist == [int(c) for c in o.c_hist[: o.n_iters_run]]
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    idx2 = NNDescent(G["X"], metric, n_neighbors=int(G["k"]), random_state=int(G["seed"]), n_jobs=int(G["T"]),
                     push_capacity=1 << 13, **kw)
    assert np.array_equal(idx2.neighbor_graph[0], gi)
    with pytest.raises(ValueError, match="Init graph size does not match dataset size!"):
        NNDescent(G["X"], metric, n_neighbors=10, random_state=1, init_graph=G["g_knn"][:-1])
