"""Host side of prepare() (pynndescent_b200/prepare.py) in the CPU suite.

`search_graph_bookkeeping` (the sparse-matrix part of _init_search_graph, pynndescent_.py:1521-1651) is fed the two
diversification passes' outputs -- computed here by the oracle on the reference's own built graph -- and must return the
reference's search graph, leaf order and min_distance (query goldens, tests/golden/make_golden_query.py).
`flatten_hub_tree` (rp_trees.py:2926-3049 convert_tree_format) is checked against a plain recursive restatement on random
level-ordered trees (the node table layout nnd_hub_tree_build produces).
"""
import numpy as np
import pytest

from conftest import QUERY_GOLDENS, load_golden


def _diversify_passes(orc, G, packed=False):
    """(di, dd, di2) exactly as prepare_search_structures obtains them from the GPU: forward diversify, zero -> eps, and the
    second pass restated literally as in the oracle (diversify_csr on the forward rows, pynndescent_.py:1552-1587)."""
    import scipy.sparse as sp
    X = np.ascontiguousarray(G["X"], np.float32); n, d = X.shape
    metric = str(G["metric"]); mid = orc.METRIC_IDS[metric]
    di = np.ascontiguousarray(G["ng_idx"], np.int32).copy(); dd = np.ascontiguousarray(G["ng_dist"], np.float32).copy()
    orc.lib().orc_diversify(di, dd, n, di.shape[1], X, d, mid)
    dd[dd == 0.0] = np.finfo(np.float32).eps
    g = sp.coo_array((n, n), dtype=np.float32)
    g.row = np.repeat(np.arange(n, dtype=np.int32), di.shape[1]); g.col = di.ravel(); g.data = dd.ravel()
    g = g.tocsr(); g.data[g.indices == -1] = 0.0; g.eliminate_zeros()
    gdata = np.ascontiguousarray(g.data, np.float32)
    orc.lib().orc_diversify_csr(np.ascontiguousarray(g.indptr, np.int32), np.ascontiguousarray(g.indices, np.int32), gdata, n, X, d, mid)
    di2 = di.copy()
    cnt = np.diff(g.indptr)
    assert np.array_equal(cnt, (di >= 0).sum(1))                      # diversified rows are left-packed: CSR entry j <-> column j
    col = np.arange(di.shape[1])[None, :]
    dropped = np.zeros(di.shape, bool)
    dropped[col < cnt[:, None]] = gdata == 0.0
    di2[dropped] = -1
    dd2 = dd.copy()
    if packed:        # what nnd_diversify hands back: every pass packs its own survivors to the left, distances alongside
        order = np.argsort(di2 < 0, axis=1, kind="stable")
        di2 = np.take_along_axis(di2, order, 1); dd2 = np.take_along_axis(dd2, order, 1)
        dd2[di2 < 0] = np.inf
    return di, dd, di2, dd2


@pytest.mark.parametrize("packed", [False, True])
@pytest.mark.parametrize("name", list(QUERY_GOLDENS))
def test_search_graph_bookkeeping_matches_reference(name, packed):
    from oracle import nnd_oracle as orc
    from pynndescent_b200.prepare import search_graph_bookkeeping
    G = load_golden(name)
    di, dd, di2, dd2 = _diversify_passes(orc, G, packed)
    S = search_graph_bookkeeping(di, dd, di2, dd2, np.ascontiguousarray(G["vertex_order"], np.int32), int(G["n_neighbors"]))
    for key in ("indptr", "indices", "tree_indices", "vertex_order"):
        assert S[key].dtype == np.int32 and np.array_equal(S[key], G[key]), key
    assert np.float32(S["min_distance"]) == np.float32(G["min_distance"])


def test_degree_prune_ties_and_hubs():
    """Rows far above max_degree, ties at the cut (kept, utils: `row[row > cut] = 0`), and rows exactly at max_degree."""
    from pynndescent_b200.prepare import search_graph_bookkeeping
    rs = np.random.RandomState(0)
    n, k, nn = 400, 12, 4                                         # max_degree = round(1.5 * 4) = 6
    di = np.full((n, k), -1, np.int32); dd = np.full((n, k), np.inf, np.float32)
    for i in range(n):
        m = rs.randint(0, k + 1)
        nb = rs.choice(np.delete(np.arange(n), i), m, replace=False)
        nb[: m // 2] = rs.choice(20, m // 2)                      # the first 20 vertices become hubs (some duplicates -> dropped)
        nb = np.array(list(dict.fromkeys(int(v) for v in nb if v != i)), np.int32)
        w = np.sort(np.round(rs.random_sample(len(nb)), 1).astype(np.float32) + np.float32(0.1))   # coarse weights -> ties
        di[i, : len(nb)] = nb; dd[i, : len(nb)] = w
    tree_idx = rs.permutation(n).astype(np.int32)
    S = search_graph_bookkeeping(di, dd, di.copy(), dd.copy(), tree_idx, nn)
    # the reference's own sequence (scipy, row by row)
    import scipy.sparse as sp
    g = sp.coo_array((n, n), dtype=np.float32)
    g.row = np.repeat(np.arange(n, dtype=np.int32), k); g.col = di.ravel(); g.data = np.where(di >= 0, dd, 0).ravel().astype(np.float32)
    g = g.tocsr(); g.data[g.indices == -1] = 0.0; g.eliminate_zeros()
    g = g.maximum(g.transpose().tocsr()).tocsr(); g.setdiag(0.0); g.eliminate_zeros()
    for i in range(n):
        row = g.data[g.indptr[i]: g.indptr[i + 1]]
        if row.shape[0] > 6:
            row[row > np.sort(row)[6]] = 0.0
    g.eliminate_zeros()
    assert np.diff(g.indptr).max() > 8                            # hubs really exist, and ties keep some rows above the cap
    g = (g != 0).astype(np.uint8)
    g = g[tree_idx, :].tocsc()[:, tree_idx].tocsr(); g.sort_indices()
    assert np.array_equal(S["indptr"], g.indptr) and np.array_equal(S["indices"], g.indices)


def test_second_pass_survivors_need_not_be_position_aligned():
    """nnd_diversify packs its survivors, so (di2, dd2) is a re-packed subset of (di, dd): same graph either way.  (In all
    five reference goldens the second pass drops nothing, so only synthetic drops exercise this.)"""
    from pynndescent_b200.prepare import search_graph_bookkeeping
    rs = np.random.RandomState(3)
    n, k = 600, 10
    di = np.full((n, k), -1, np.int32); dd = np.full((n, k), np.inf, np.float32)
    for i in range(n):
        m = rs.randint(2, k + 1)
        di[i, :m] = rs.choice(np.delete(np.arange(n), i), m, replace=False)
        dd[i, :m] = np.sort(rs.random_sample(m).astype(np.float32) + np.float32(0.05))
    di2 = di.copy(); di2[(rs.random_sample(di.shape) < 0.3) & (di >= 0)] = -1
    order = np.argsort(di2 < 0, axis=1, kind="stable")
    pi = np.take_along_axis(di2, order, 1); pd = np.take_along_axis(dd, order, 1); pd[pi < 0] = np.inf
    assert not np.array_equal(pi, di2)
    vo = rs.permutation(n).astype(np.int32)
    A = search_graph_bookkeeping(di, dd, di2, dd, vo, 4)
    B = search_graph_bookkeeping(di, dd, pi, pd, vo, 4)
    assert np.array_equal(A["indptr"], B["indptr"]) and np.array_equal(A["indices"], B["indices"])
    assert A["min_distance"] == B["min_distance"] == dd[di >= 0].min()


def _flatten_reference(start, size, left, right, hrow, hyp, off, d):
    """convert_tree_format restated recursively (rp_trees.py:2926-3049): pre-order numbering, leaves encode -start / -end."""
    n_nodes = len(start)
    hyperplanes = np.zeros((n_nodes, d), np.float32); offsets = np.zeros(n_nodes, np.float32)
    children = np.full((n_nodes, 2), -1, np.int32)
    counter = [0]

    def rec(v):
        me = counter[0]; counter[0] += 1
        if left[v] >= 0:
            hyperplanes[me] = hyp[hrow[v]]; offsets[me] = off[hrow[v]]
            children[me, 0] = rec(left[v]); children[me, 1] = rec(right[v])
        else:
            children[me, 0] = -start[v]; children[me, 1] = -(start[v] + size[v])
        return me
    import sys
    sys.setrecursionlimit(100000)
    rec(0)
    return hyperplanes, offsets, children


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_flatten_hub_tree_matches_recursive_restatement(seed):
    from pynndescent_b200.prepare import flatten_hub_tree
    rs = np.random.RandomState(seed)
    # random binary tree whose nodes are numbered level by level (parents before children), like nnd_hub_tree_build's table
    left, right, start, size, hrow = [-1], [-1], [0], [1000], [-1]
    frontier = [0]; nh = 0
    while frontier:
        nxt = []
        for v in frontier:
            if size[v] > 12 and rs.random_sample() < 0.9:
                sl = int(rs.randint(1, size[v])); hrow[v] = nh; nh += 1
                for s0, sz in ((start[v], sl), (start[v] + sl, size[v] - sl)):
                    left.append(-1); right.append(-1); start.append(s0); size.append(sz); hrow.append(-1); nxt.append(len(left) - 1)
                left[v], right[v] = len(left) - 2, len(left) - 1
        frontier = nxt
    a = [np.asarray(x, np.int32) for x in (start, size, left, right, hrow)]
    d = 5
    hyp = rs.standard_normal((nh, d)).astype(np.float32); off = rs.standard_normal(nh).astype(np.float32)
    got = flatten_hub_tree(*a, hyp, off, d)
    exp = _flatten_reference(*a, hyp, off, d)
    for g_, e_ in zip(got, exp):
        assert np.array_equal(g_, e_)
