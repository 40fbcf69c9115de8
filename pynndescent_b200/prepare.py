"""Search-graph preparation (pynndescent_.py:1333-1662 _init_search_graph, dense data, standard diversification,
diversify_prob = 1 -- the defaults).

GPU (libnnd_b200.so): the hub search tree (rp_trees.py:1233 make_hub_tree) and both diversification passes
(all the distance work).  Host: the sparse-matrix bookkeeping, done with the very scipy calls the reference makes
(coo -> csr, eliminate_zeros, transpose, maximum, setdiag, degree prune, leaf-order permutation) and the pre-order
flattening of the tree (rp_trees.py:2926-3049 convert_tree_format).
"""
import ctypes as C

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
    # pre-order numbering (rp_trees.py:2926-2961): left child = me + 1, right child = me + 1 + |left subtree|
    sub = np.ones(n_nodes, np.int64)
    order = []
    stack = [0]
    while stack:                       # children are created after their parent, so ids ascend with depth
        v = stack.pop(); order.append(v)
        if left[v] >= 0:
            stack.append(right[v]); stack.append(left[v])
    for v in reversed(order):
        if left[v] >= 0:
            sub[v] = 1 + sub[left[v]] + sub[right[v]]
    num = np.zeros(n_nodes, np.int64)
    for v in order:                    # `order` is a pre-order walk
        if left[v] >= 0:
            num[left[v]] = num[v] + 1
            num[right[v]] = num[v] + 1 + sub[left[v]]
    hyperplanes = np.zeros((n_nodes, d), np.float32); offsets = np.zeros(n_nodes, np.float32)
    children = np.full((n_nodes, 2), -1, np.int32)
    for v in order:
        me = num[v]
        if left[v] >= 0:
            hyperplanes[me] = hyp[hrow[v]]; offsets[me] = off[hrow[v]]
            children[me, 0] = num[left[v]]; children[me, 1] = num[right[v]]
        else:
            children[me, 0] = -start[v]; children[me, 1] = -(start[v] + size[v])
    return hyperplanes, offsets, children, indices.cpu().numpy()


def prepare_search_structures(eng, metric, ng_idx_t, ng_dist_t, n_neighbors, angular, leaf_size=None,
                              pruning_degree_multiplier=1.5, max_depth=200):
    """ng_idx_t / ng_dist_t: the deheap-sorted graph on the device with RAW (alternative-metric) distances."""
    import scipy.sparse as sp
    n, k = eng.N, ng_idx_t.shape[1]
    search_leaf_size = leaf_size if leaf_size is not None else 30                              # :1341-1345
    hyperplanes, offsets, children, tree_idx = build_hub_tree(eng, ng_idx_t, search_leaf_size, angular, max_depth)
    di_t = ng_idx_t.clone(); dd_t = ng_dist_t.clone()
    P = eng.problem()
    check(lib.nnd_diversify(eng._ctx, C.byref(P), _p(di_t), _p(dd_t), eng._st()))            # :1502-1519
    # second pass == diversify_csr on the forward rows (:1552-1587: the "reverse" graph object shares its arrays with
    # the forward CSR, see oracle/nnd_oracle.py prepare()); zero distances are FLOAT32_EPS by then (:1526)
    dd_t[dd_t == 0.0] = float(FLOAT32_EPS)
    di2_t = di_t.clone(); dd2_t = dd_t.clone()
    check(lib.nnd_diversify(eng._ctx, C.byref(P), _p(di2_t), _p(dd2_t), eng._st()))
    eng.stream.synchronize()
    di = di_t.cpu().numpy(); dd = dd_t.cpu().numpy()
    g = sp.coo_array((n, n), dtype=np.float32)                                                  # :1521-1537
    g.row = np.repeat(np.arange(n, dtype=np.int32), k); g.col = di.ravel(); g.data = dd.ravel()
    g = g.tocsr()
    g.data[g.indices == -1] = 0.0
    g.eliminate_zeros()
    min_distance = np.min(g.data)                                                               # :1539
    di2 = di2_t.cpu().numpy()
    if not np.array_equal(di2, di):                                                             # entries the second pass dropped
        keep = np.zeros(g.data.shape[0], bool)
        for i in np.where((di2 != di).any(1))[0]:
            cols = g.indices[g.indptr[i]: g.indptr[i + 1]]
            keep_i = np.isin(cols, di2[i][di2[i] >= 0])
            g.data[g.indptr[i]: g.indptr[i + 1]][~keep_i] = 0.0
        g.eliminate_zeros()
    rev = g.transpose().tocsr()
    rev.sort_indices(); g.sort_indices()
    g = g.maximum(rev).tocsr()                                                                  # :1601
    g.setdiag(0.0); g.eliminate_zeros()
    max_degree = int(np.round(pruning_degree_multiplier * n_neighbors))                         # :1608-1611, :728-760
    deg = np.diff(g.indptr)
    for i in np.where(deg > max_degree)[0]:
        row = g.data[g.indptr[i]: g.indptr[i + 1]]
        cut = np.sort(row)[max_degree]
        row[row > cut] = 0.0
    g.eliminate_zeros()
    g = (g != 0).astype(np.uint8)                                                               # :1613
    vertex_order = tree_idx                                                                     # :1630
    g = g[vertex_order, :].tocsc()[:, vertex_order].tocsr(); g.sort_indices()                   # :1631-1634
    tree_order = np.argsort(vertex_order)
    new_tree_idx = tree_idx[tree_order].astype(np.int32)                                        # :763-771
    return dict(indptr=g.indptr.astype(np.int32), indices=g.indices.astype(np.int32), hyperplanes=hyperplanes, offsets=offsets,
                children=children, tree_indices=new_tree_idx, vertex_order=vertex_order.astype(np.int32),
                min_distance=np.float32(min_distance))
