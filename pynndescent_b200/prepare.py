"""Search-graph preparation (pynndescent_.py:1333-1662 _init_search_graph, dense data, standard diversification,
diversify_prob = 1 -- the defaults).

GPU (libnnd_b200.so): the hub search tree (rp_trees.py:1233 make_hub_tree) and both diversification passes
(all the distance work).  Host: the sparse-matrix bookkeeping (`search_graph_bookkeeping`: the reference's coo -> csr,
transpose, maximum, setdiag, degree prune, leaf-order permutation, restated without per-row Python loops) and the
pre-order flattening of the tree (`flatten_hub_tree`, rp_trees.py:2926-3049 convert_tree_format), both tested against
the reference's structures in the CPU suite (tests/test_prepare_host.py).
"""
import ctypes as C
import os

import numpy as np
import torch

from ._ffi import check, lib

FLOAT32_EPS = np.finfo(np.float32).eps


def _p(t):
    return C.c_void_p(t.data_ptr())


def build_hub_tree(eng, ng_idx_t, leaf_size, angular, max_depth):
    """-> (hyperplanes [n_nodes,d], offsets, children [n_nodes,2], indices [N]) in the reference's flat-tree format."""
    N, d, ld, dev = eng.N, eng.d, eng.ld, eng.device
    nb = C.c_size_t()
    check(lib.nnd_hub_tree_workspace_bytes(N, d, C.byref(nb)))
    ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
    indices = torch.empty(N, dtype=torch.int32, device=dev)
    ncap = 2 * N + 16
    n_start = torch.empty(ncap, dtype=torch.int32, device=dev); n_size = torch.empty(ncap, dtype=torch.int32, device=dev)
    n_left = torch.empty(ncap, dtype=torch.int32, device=dev); n_right = torch.empty(ncap, dtype=torch.int32, device=dev)
    n_hrow = torch.empty(ncap, dtype=torch.int32, device=dev)
    hyper_cap = N // 8 + 64
    n_hyper = torch.empty((hyper_cap, ld), dtype=torch.float32, device=dev); n_off = torch.empty(hyper_cap, dtype=torch.float32, device=dev)
    nn = C.c_int32(); nh = C.c_int32()
    check(lib.nnd_hub_tree_build(eng._ctx, _p(eng.X), N, d, ld, int(bool(angular)), _p(ng_idx_t), ng_idx_t.shape[1], int(leaf_size),
                                 int(max_depth), _p(indices), _p(n_start), _p(n_size), _p(n_left), _p(n_right), _p(n_hrow),
                                 _p(n_hyper), _p(n_off), hyper_cap, C.byref(nn), C.byref(nh), _p(ws), nb.value, eng._st()))
    n_nodes = nn.value
    start = n_start[:n_nodes].cpu().numpy(); size = n_size[:n_nodes].cpu().numpy()
    left = n_left[:n_nodes].cpu().numpy(); right = n_right[:n_nodes].cpu().numpy(); hrow = n_hrow[:n_nodes].cpu().numpy()
    hyp = n_hyper[: nh.value, :d].cpu().numpy(); off = n_off[: nh.value].cpu().numpy()
    hyperplanes, offsets, children = flatten_hub_tree(start, size, left, right, hrow, hyp, off, d)
    return hyperplanes, offsets, children, indices.cpu().numpy(), indices


def flatten_hub_tree(start, size, left, right, hrow, hyp, off, d):
    """Node table of nnd_hub_tree_build (nodes numbered level by level, parents before children; left/right = -1 for a
    leaf, hrow = row of the node's hyperplane) -> the reference's flat pre-order arrays (rp_trees.py:2926-3049
    convert_tree_format): node `me` has children me + 1 and me + 1 + |left subtree|, a leaf stores (-start, -end) of its
    slice of `indices`.  Level-synchronous like the build: subtree sizes bottom-up, pre-order numbers top-down, one
    vectorised step per tree level instead of a Python visit per node."""
    n_nodes = len(start)
    left = np.asarray(left, np.int64); right = np.asarray(right, np.int64)
    inner = left >= 0
    depth = np.zeros(n_nodes, np.int64)
    levels = [np.array([0], np.int64)]
    while True:                                        # children of level L form level L + 1
        cur = levels[-1]; cur = cur[inner[cur]]
        if cur.size == 0:
            break
        nxt = np.concatenate([left[cur], right[cur]])
        depth[nxt] = len(levels)
        levels.append(nxt)
    sub = np.ones(n_nodes, np.int64)
    for lv in reversed(levels):
        v = lv[inner[lv]]
        sub[v] = 1 + sub[left[v]] + sub[right[v]]
    num = np.zeros(n_nodes, np.int64)
    for lv in levels:
        v = lv[inner[lv]]
        num[left[v]] = num[v] + 1
        num[right[v]] = num[v] + 1 + sub[left[v]]
    hyperplanes = np.zeros((n_nodes, d), np.float32); offsets = np.zeros(n_nodes, np.float32)
    children = np.full((n_nodes, 2), -1, np.int32)
    vi = np.flatnonzero(inner); vl = np.flatnonzero(~inner)
    hyperplanes[num[vi]] = np.asarray(hyp)[np.asarray(hrow)[vi], :d]; offsets[num[vi]] = np.asarray(off)[np.asarray(hrow)[vi]]
    children[num[vi], 0] = num[left[vi]]; children[num[vi], 1] = num[right[vi]]
    children[num[vl], 0] = -np.asarray(start)[vl]; children[num[vl], 1] = -(np.asarray(start)[vl] + np.asarray(size)[vl])
    return hyperplanes, offsets, children


def prepare_search_structures(eng, metric, ng_idx_t, ng_dist_t, n_neighbors, angular, leaf_size=None,
                              pruning_degree_multiplier=1.5, max_depth=200):
    """ng_idx_t / ng_dist_t: the deheap-sorted graph on the device with RAW (alternative-metric) distances."""
    search_leaf_size = leaf_size if leaf_size is not None else 30                              # :1341-1345
    hyperplanes, offsets, children, tree_idx, tree_idx_t = build_hub_tree(eng, ng_idx_t, search_leaf_size, angular, max_depth)
    di_t = ng_idx_t.clone(); dd_t = ng_dist_t.clone()
    P = eng.problem()
    check(lib.nnd_diversify(eng._ctx, C.byref(P), _p(di_t), _p(dd_t), eng._st()))            # :1502-1519
    # second pass == diversify_csr on the forward rows (:1552-1587: the "reverse" graph object shares its arrays with
    # the forward CSR, see oracle/nnd_oracle.py prepare()); zero distances are FLOAT32_EPS by then (:1526)
    dd_t[dd_t == 0.0] = float(FLOAT32_EPS)
    di2_t = di_t.clone(); dd2_t = dd_t.clone()
    check(lib.nnd_diversify(eng._ctx, C.byref(P), _p(di2_t), _p(dd2_t), eng._st()))
    if os.environ.get("NND_PREPARE_GPU", "1") != "0":    # default: the bookkeeping on the device (k_search_graph.cu); "0" = host scipy
        S = search_graph_device(eng, di_t, dd_t, di2_t, dd2_t, tree_idx_t, tree_idx, n_neighbors, pruning_degree_multiplier)
    else:
        eng.stream.synchronize()
        S = search_graph_bookkeeping(di_t.cpu().numpy(), dd_t.cpu().numpy(), di2_t.cpu().numpy(), dd2_t.cpu().numpy(), tree_idx,
                                     n_neighbors, pruning_degree_multiplier)
    S.update(hyperplanes=hyperplanes, offsets=offsets, children=children)
    return S


def search_graph_device(eng, di_t, dd_t, di2_t, dd2_t, vertex_order_t, vertex_order, n_neighbors, pruning_degree_multiplier=1.5):
    """search_graph_bookkeeping on the device (nnd_search_graph_build): only the finished CSR comes back to the host."""
    n, k = di_t.shape
    nb = C.c_size_t()
    check(lib.nnd_search_graph_workspace_bytes(n, k, C.byref(nb)))
    ws = torch.empty(nb.value, dtype=torch.uint8, device=eng.device)
    cap = 2 * n * k
    indptr = torch.empty(n + 1, dtype=torch.int32, device=eng.device)
    indices = torch.empty(cap, dtype=torch.int32, device=eng.device)
    nnz = C.c_int64(); md = C.c_float()
    max_degree = int(np.round(pruning_degree_multiplier * n_neighbors))
    check(lib.nnd_search_graph_build(eng._ctx, _p(di_t), _p(dd_t), _p(di2_t), _p(dd2_t), n, k, _p(vertex_order_t), max_degree, _p(indptr),
                                     _p(indices), cap, C.byref(nnz), C.byref(md), _p(ws), nb.value, eng._st()))
    eng.stream.synchronize()
    return dict(indptr=indptr.cpu().numpy(), indices=indices[: nnz.value].cpu().numpy(),
                tree_indices=np.arange(n, dtype=np.int32),                      # :763-771 resort_tree_indices of a permutation
                vertex_order=np.asarray(vertex_order, np.int32), min_distance=np.float32(md.value))


def search_graph_bookkeeping(di, dd, di2, dd2, tree_idx, n_neighbors, pruning_degree_multiplier=1.5):
    """The sparse-matrix part of _init_search_graph (pynndescent_.py:1521-1651) on the host, from the two diversification
    passes' outputs: (`di`, `dd`) = forward-diversified rows (-1 holes, zero distances already FLOAT32_EPS, :1526) and
    (`di2`, `dd2`) = the same rows after the second (CSR) pass: a subset of the first pass's entries with the same distances,
    but PACKED by the pass itself, so the two pairs are not position-aligned.

    Same result as the reference's scipy sequence (coo -> csr, eliminate_zeros, transpose, maximum, setdiag(0), degree
    prune, != 0, leaf-order permutation), restated so that nothing loops over rows in Python and no sparsity-changing
    scipy call is used where a mask does (`setdiag` alone cost more than everything else together at 1M points):
    the first-pass matrix only contributes `min_distance` (:1539), the symmetrised matrix is built straight from the
    second-pass survivors, the diagonal and the degree prune are masks on the CSR arrays.
    """
    import scipy.sparse as sp
    n, k = di.shape
    min_distance = np.min(dd[di >= 0])                                                           # :1539 (first-pass graph)
    keep = di2 >= 0
    indptr = np.zeros(n + 1, np.int64); np.cumsum(keep.sum(1), out=indptr[1:])                 # rows are already grouped and
    g = sp.csr_array((dd2[keep], di2[keep], indptr), shape=(n, n))                             # duplicate-free: CSR as is
    g.eliminate_zeros()                                                                         # :1521-1537 + :1552-1587
    rev = g.transpose().tocsr()
    rev.sort_indices(); g.sort_indices()
    g = g.maximum(rev).tocsr()                                                                  # :1601
    row_of = np.repeat(np.arange(n, dtype=np.int32), np.diff(g.indptr))
    g.data[g.indices == row_of] = 0.0                                                           # :1605 setdiag(0)
    g.eliminate_zeros()
    max_degree = int(np.round(pruning_degree_multiplier * n_neighbors))                         # :1608-1611, :728-760
    deg = np.diff(g.indptr)
    hub = np.flatnonzero(deg > max_degree)
    if hub.size:
        # degree_prune: in every row with more than max_degree entries, zero what is larger than the (max_degree+1)-th
        # smallest weight (ties with it stay) -- all hub rows at once: sort their entries by (row, weight)
        seg = np.zeros(n, bool); seg[hub] = True
        sel = np.flatnonzero(seg[np.repeat(np.arange(n, dtype=np.int32), deg)])
        r = np.repeat(np.arange(n, dtype=np.int32), deg)[sel]; w = g.data[sel]
        order = np.lexsort((w, r))
        first = np.searchsorted(r[order], hub)                                                  # start of every hub row in `order`
        cut = np.empty(n, np.float32); cut[hub] = w[order][first + max_degree]
        g.data[sel[w > cut[r]]] = 0.0
        g.eliminate_zeros()
    vertex_order = tree_idx                                                                     # :1630
    # :1613 + :1631-1634  B[a, b] = (g[vo[a], vo[b]] != 0): new row a is old row vo[a] with its columns renamed by vo^-1
    inv = np.empty(n, np.int32); inv[vertex_order] = np.arange(n, dtype=np.int32)
    deg = np.diff(g.indptr)[vertex_order]
    new_indptr = np.zeros(n + 1, np.int64); np.cumsum(deg, out=new_indptr[1:])
    it = np.int32 if g.indptr[-1] < 2 ** 31 - 1 else np.int64
    src = np.repeat((g.indptr[:-1][vertex_order] - new_indptr[:-1]).astype(it), deg) + np.arange(new_indptr[-1], dtype=it)
    h = sp.csr_array((np.ones(src.shape[0], np.uint8), inv[g.indices[src]], new_indptr), shape=(n, n))
    h.sort_indices()
    new_tree_idx = tree_idx[np.argsort(vertex_order)].astype(np.int32)                          # :763-771
    return dict(indptr=h.indptr.astype(np.int32), indices=h.indices.astype(np.int32), tree_indices=new_tree_idx,
                vertex_order=np.asarray(vertex_order, np.int32), min_distance=np.float32(min_distance))
