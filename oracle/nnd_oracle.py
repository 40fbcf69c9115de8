"""ctypes front-end of the CPU oracle (oracle/nnd_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; nothing under pynndescent_b200/ does.

`OracleNNDescent` replays the reference constructor's host logic
(/root/reference/pynndescent/pynndescent_.py:976-1269) on top of the C restatement so
that every intermediate (leaf_array, heap after init, candidates, ...) can be dumped and
fed to the CUDA stages.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

METRIC_IDS = {"sqeuclidean": 0, "euclidean": 0, "l2": 0, "cosine": 1, "dot": 2, "correlation": 3,
              # SURVEY 8f-2 (ids restate include/nnd_b200.h)
              "manhattan": 4, "taxicab": 4, "l1": 4, "chebyshev": 5, "linfinity": 5, "linfty": 5, "linf": 5,
              "canberra": 6, "hamming": 7, "jaccard": 8, "dice": 9, "matching": 10, "kulsinski": 11,
              "rogerstanimoto": 12, "sokalmichener": 12, "russellrao": 13, "sokalsneath": 14, "yule": 15,
              "hellinger": 16,
              # alternative_inner_product (distances.py:759-790); true_angular computes in alternative_cosine (distances.py:2181-2184)
              "inner_product": 17, "true_angular": 1,
              # distances.py:138-153 minkowski(x, y, p): p is process-wide state of the oracle (set_metric_param)
              "minkowski": 18}
ANGULAR_METRICS = ("cosine", "dot", "correlation", "dice", "jaccard", "hellinger", "hamming")
INT32_MIN = np.iinfo(np.int32).min + 1
INT32_MAX = np.iinfo(np.int32).max - 1

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(force=False):
    so = os.path.join(_HERE, "libnnd_oracle.so")
    src = os.path.join(_HERE, "nnd_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], env={**os.environ, "CC": "gcc"})
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    L.orc_tau_rand_int.argtypes = [_i64p]; L.orc_tau_rand_int.restype = C.c_int32
    L.orc_tau_rand.argtypes = [_i64p]; L.orc_tau_rand.restype = C.c_float
    L.orc_det_log2.argtypes = [C.c_double]; L.orc_det_log2.restype = C.c_double
    L.orc_dist.argtypes = [C.c_int, _f32p, _f32p, C.c_int]; L.orc_dist.restype = C.c_float
    L.orc_checked_flagged_heap_push.argtypes = [_f32p, _i32p, _u8p, C.c_int, C.c_float, C.c_int32, C.c_uint8]
    L.orc_checked_flagged_heap_push.restype = C.c_int
    L.orc_checked_heap_push.argtypes = [_f32p, _i32p, C.c_int, C.c_float, C.c_int32]
    L.orc_checked_heap_push.restype = C.c_int
    L.orc_simple_heap_push.argtypes = [_f32p, _i32p, C.c_int, C.c_float, C.c_int32]
    L.orc_simple_heap_push.restype = C.c_int
    L.orc_deheap_sort.argtypes = [_i32p, _f32p, C.c_int64, C.c_int]
    L.orc_new_build_candidates.argtypes = [_i32p, _u8p, C.c_int64, C.c_int, C.c_int, _i64p, C.c_int, _i32p, _i32p]
    L.orc_process_candidates.argtypes = [_f32p, C.c_int64, C.c_int, C.c_int, _i32p, _f32p, _u8p, C.c_int, _i32p,
                                         _i32p, C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_int64)]
    L.orc_process_candidates.restype = C.c_int64
    L.orc_generate_updates.argtypes = [_f32p, C.c_int64, C.c_int, C.c_int, _i32p, _i32p, C.c_int, _f32p, C.c_int64,
                                       C.c_int64, _i32p, _i32p, _f32p, C.c_int64]
    L.orc_generate_updates.restype = C.c_int64
    L.orc_init_rp_tree.argtypes = [_f32p, C.c_int64, C.c_int, C.c_int, _i32p, _f32p, _u8p, C.c_int, _i32p, C.c_int64,
                                   C.c_int]
    L.orc_init_random.argtypes = [_f32p, C.c_int64, C.c_int, C.c_int, _i32p, _f32p, _u8p, C.c_int, _i64p]
    L.orc_init_from_graph.argtypes = [_f32p, C.c_int64, C.c_int, C.c_int, _i32p, _f32p, _u8p, C.c_int, _i32p, C.c_int,
                                      C.c_void_p]
    L.orc_init_from_neighbor_graph.argtypes = [_i32p, _f32p, _u8p, C.c_int, C.c_int64, _i32p, _f32p, C.c_int]
    L.orc_make_forest.argtypes = [_f32p, C.c_int64, C.c_int, C.c_int, _i64p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.orc_make_forest.restype = C.c_void_p
    L.orc_forest_free.argtypes = [C.c_void_p]
    L.orc_forest_leaf_dims.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int)]
    L.orc_forest_leaf_array.argtypes = [C.c_void_p, C.c_int, _i32p]
    L.orc_forest_tree_nodes.argtypes = [C.c_void_p, C.c_int]; L.orc_forest_tree_nodes.restype = C.c_int64
    L.orc_forest_tree_arrays.argtypes = [C.c_void_p, C.c_int, _i32p, _i32p, _f32p, _f32p, _i32p]
    L.orc_nn_descent_internal.argtypes = [_f32p, C.c_int64, C.c_int, C.c_int, _i32p, _f32p, _u8p, C.c_int, C.c_int,
                                          _i64p, C.c_int, C.c_int, C.c_double, C.c_int64, _i64p]
    L.orc_nn_descent_internal.restype = C.c_int
    L.orc_search.argtypes = [_f32p, C.c_int64, C.c_int, C.c_int, _i32p, _i32p, _f32p, _f32p, _i32p, _i32p, _f32p, C.c_int64,
                             C.c_int, C.c_double, C.c_float, C.c_int, C.c_int, _i64p, _i32p, _f32p]
    L.orc_hub_tree.argtypes = [_f32p, C.c_int64, C.c_int, C.c_int, _i32p, C.c_int, C.c_int, C.c_int, _i64p, C.POINTER(C.c_int64)]
    L.orc_hub_tree.restype = C.c_void_p
    L.orc_hub_tree_flatten.argtypes = [C.c_void_p, C.c_int, C.c_int64, _f32p, _f32p, _i32p, _i32p]
    L.orc_diversify.argtypes = [_i32p, _f32p, C.c_int64, C.c_int, _f32p, C.c_int, C.c_int]
    L.orc_diversify_csr.argtypes = [_i32p, _i32p, _f32p, C.c_int64, _f32p, C.c_int, C.c_int]
    L.orc_num_threads.restype = C.c_int
    L.orc_set_num_threads.argtypes = [C.c_int]
    L.orc_set_metric_param.argtypes = [C.c_double]
    L.orc_set_metric_param.restype = None
    _LIB = L
    return L


# ----------------------------------------------------------------- thin helpers
def tau_rand_int(state):
    return int(lib().orc_tau_rand_int(state))


def tau_rand(state):
    return float(lib().orc_tau_rand(state))


def set_metric_param(metric, metric_kwds=None):
    """metric_kwds.values() are passed positionally to the distance function (pynndescent_.py:1064-1065, 1294-1302); the only
    parametrised metric restated here is minkowski (p, default 2: distances.py:138)."""
    if metric == "minkowski":
        vals = list((metric_kwds or {}).values())
        lib().orc_set_metric_param(float(vals[0]) if vals else 2.0)
    elif metric_kwds:
        raise NotImplementedError("metric_kwds of " + metric)


def dist(metric, x, y):
    x = np.ascontiguousarray(x, np.float32); y = np.ascontiguousarray(y, np.float32)
    return float(lib().orc_dist(METRIC_IDS[metric], x, y, x.shape[0]))


def make_heap(n, k):
    return (np.full((n, k), -1, np.int32), np.full((n, k), np.inf, np.float32), np.zeros((n, k), np.uint8))


def make_forest(X, n_neighbors, n_trees, leaf_size, random_state, angular=False, max_depth=200, scheme="dfs"):
    """rp_trees.py:2815-2922 make_forest + rptree_leaf_array -> (leaf_array, rng_states)."""
    if leaf_size is None:
        leaf_size = max(60, min(256, 5 * int(n_neighbors)))
    rng_states = random_state.randint(INT32_MIN, INT32_MAX, size=(n_trees, 3)).astype(np.int64)
    return forest_leaf_array(X, rng_states, leaf_size, angular, max_depth, scheme), rng_states


def forest_leaf_array(X, rng_states, leaf_size, angular=False, max_depth=200, scheme="dfs"):
    L = lib()
    n_trees = rng_states.shape[0]
    if n_trees == 0:
        return np.array([[-1]], np.int32)
    f = L.orc_make_forest(X, X.shape[0], X.shape[1], int(angular), np.ascontiguousarray(rng_states), n_trees,
                          int(leaf_size), int(max_depth), 0 if scheme == "dfs" else 1)
    nl = C.c_int64(); ml = C.c_int()
    L.orc_forest_leaf_dims(f, int(leaf_size), C.byref(nl), C.byref(ml))
    out = np.empty((nl.value, ml.value), np.int32)
    L.orc_forest_leaf_array(f, ml.value, out)
    L.orc_forest_free(f)
    return out


def init_rp_tree(X, metric, heap, leaf_array):
    la = np.ascontiguousarray(leaf_array, np.int32)
    lib().orc_init_rp_tree(X, X.shape[0], X.shape[1], METRIC_IDS[metric], heap[0], heap[1], heap[2], heap[0].shape[1],
                           la, la.shape[0], la.shape[1])


def init_random(X, metric, heap, rng_state):
    lib().orc_init_random(X, X.shape[0], X.shape[1], METRIC_IDS[metric], heap[0], heap[1], heap[2], heap[0].shape[1],
                          rng_state)


def init_from_graph(X, metric, heap, init_graph, init_dist=None):
    """utils.py:836-860 warm start (distances from `metric`, or the caller's init_dist pushed verbatim)."""
    g = np.ascontiguousarray(init_graph, np.int32)
    dd = None if init_dist is None else np.ascontiguousarray(init_dist, np.float32)
    lib().orc_init_from_graph(X, X.shape[0], X.shape[1], METRIC_IDS[metric], heap[0], heap[1], heap[2], heap[0].shape[1],
                              g, g.shape[1], None if dd is None else dd.ctypes.data_as(C.c_void_p))


def init_from_neighbor_graph(heap, indices, distances):
    """pynndescent_.py:207-214 (update()): unflagged pushes of the old graph's rows into the first rows of `heap`."""
    g = np.ascontiguousarray(indices, np.int32); dd = np.ascontiguousarray(distances, np.float32)
    lib().orc_init_from_neighbor_graph(heap[0], heap[1], heap[2], heap[0].shape[1], g.shape[0], g, dd, g.shape[1])


def new_build_candidates(heap, max_candidates, rng_state, n_threads):
    n, k = heap[0].shape
    newc = np.empty((n, max_candidates), np.int32); oldc = np.empty((n, max_candidates), np.int32)
    lib().orc_new_build_candidates(heap[0], heap[2], n, k, max_candidates, rng_state, n_threads, newc, oldc)
    return newc, oldc


def generate_updates(X, metric, newc, oldc, thr, r0=0, r1=None):
    n, c = newc.shape
    r1 = n if r1 is None else r1
    cap = (r1 - r0) * (c * (c + 1) // 2 + c * c) + 16
    p = np.empty(cap, np.int32); q = np.empty(cap, np.int32); d = np.empty(cap, np.float32)
    m = lib().orc_generate_updates(X, n, X.shape[1], METRIC_IDS[metric], newc, oldc, c,
                                   np.ascontiguousarray(thr, np.float32), r0, r1, p, q, d, cap)
    return p[:m].copy(), q[:m].copy(), d[:m].copy()


def process_candidates(X, metric, heap, newc, oldc, n_threads, block_size=16384):
    n, k = heap[0].shape
    em = C.c_int64()
    c = lib().orc_process_candidates(X, n, X.shape[1], METRIC_IDS[metric], heap[0], heap[1], heap[2], k, newc, oldc,
                                     newc.shape[1], block_size, n_threads, C.byref(em))
    return int(c), int(em.value)


def deheap_sort(ind, dist_):
    lib().orc_deheap_sort(ind, dist_, ind.shape[0], ind.shape[1])
    return ind, dist_


def search(data, metric, indptr, indices, hyperplanes, offsets, children, tree_indices, Q, k, epsilon, min_distance,
           n_neighbors, rng_state):
    """pynndescent_.py:1793-1883 search_closure -> raw (heap_idx, heap_dist) (not yet deheap-sorted)."""
    data = np.ascontiguousarray(data, np.float32); Q = np.ascontiguousarray(Q, np.float32)
    hi = np.empty((Q.shape[0], k), np.int32); hd = np.empty((Q.shape[0], k), np.float32)
    lib().orc_search(data, data.shape[0], data.shape[1], METRIC_IDS[metric], np.ascontiguousarray(indptr, np.int32),
                     np.ascontiguousarray(indices, np.int32), np.ascontiguousarray(hyperplanes, np.float32),
                     np.ascontiguousarray(offsets, np.float32), np.ascontiguousarray(children, np.int32),
                     np.ascontiguousarray(tree_indices, np.int32), Q, Q.shape[0], int(k), float(epsilon), float(min_distance),
                     int(n_neighbors), int(metric in ("cosine", "dot")), np.ascontiguousarray(rng_state, np.int64), hi, hd)
    return hi, hd


def hub_tree(X, ng_idx, rng_state, leaf_size, angular, max_depth=200):
    """rp_trees.py:1233 make_hub_tree + :3019 convert_tree_format -> (hyperplanes, offsets, children, indices)."""
    L = lib(); n, d = X.shape
    nn = C.c_int64()
    t = L.orc_hub_tree(X, n, d, int(angular), np.ascontiguousarray(ng_idx, np.int32), ng_idx.shape[1], int(leaf_size), int(max_depth),
                       np.ascontiguousarray(rng_state, np.int64), C.byref(nn))
    hyper = np.zeros((nn.value, d), np.float32); offs = np.zeros(nn.value, np.float32)
    children = np.empty((nn.value, 2), np.int32); indices = np.empty(n, np.int32)
    L.orc_hub_tree_flatten(t, d, n, hyper, offs, children, indices)
    return hyper, offs, children, indices


def prepare(X, metric, ng_idx, ng_dist, n_neighbors, rng_state, leaf_size=None, pruning_degree_multiplier=1.5, max_depth=200):
    """pynndescent_.py:1333-1662 _init_search_graph (dense, tree_init, standard diversification, diversify_prob=1).
    Returns the structures query() walks: data (leaf order), indptr, indices, tree arrays, vertex_order, min_distance."""
    import scipy.sparse as sp
    L = lib()
    X = np.ascontiguousarray(X, np.float32); n, d = X.shape
    angular = metric in ANGULAR_METRICS
    search_leaf_size = leaf_size if leaf_size is not None else 30                        # :1341-1345
    hyper, offs, children, tree_idx = hub_tree(X, ng_idx, rng_state, search_leaf_size, angular, max_depth)
    di = np.ascontiguousarray(ng_idx, np.int32).copy(); dd = np.ascontiguousarray(ng_dist, np.float32).copy()
    L.orc_diversify(di, dd, n, di.shape[1], X, d, METRIC_IDS[metric])                  # :1502-1519
    dd[dd == 0.0] = np.finfo(np.float32).eps                                           # :1526
    g = sp.coo_array((n, n), dtype=np.float32)                                         # :1521-1537, built exactly like the
    g.row = np.repeat(np.arange(n, dtype=np.int32), di.shape[1])                       # reference does: the CSR rows keep the
    g.col = di.ravel(); g.data = dd.ravel()                                            # ascending-distance entry order
    g = g.tocsr()
    g.data[g.indices == -1] = 0.0
    g.eliminate_zeros()
    min_distance = np.min(g.data)                                                      # :1539
    # :1552-1587  `reverse_graph = self._search_graph.transpose()` is a CSC VIEW that shares indptr / indices / data with
    # the forward CSR graph, so diversify_csr walks the FORWARD rows and zeroes entries of the shared data array;
    # eliminate_zeros then compacts both.  Restated literally: prune the forward rows, then transpose.
    gdata = np.ascontiguousarray(g.data, np.float32)
    L.orc_diversify_csr(np.ascontiguousarray(g.indptr, np.int32), np.ascontiguousarray(g.indices, np.int32), gdata, n, X, d,
                        METRIC_IDS[metric])
    g = sp.csr_array((gdata, g.indices, g.indptr), shape=(n, n)); g.eliminate_zeros()
    rev = g.transpose().tocsr()
    rev.sort_indices(); g.sort_indices()
    g = g.maximum(rev).tocsr()                                                         # :1601
    g.setdiag(0.0); g.eliminate_zeros()
    max_degree = int(np.round(pruning_degree_multiplier * n_neighbors))                # :1608-1611, :728-760
    for i in range(n):
        row = g.data[g.indptr[i]: g.indptr[i + 1]]
        if row.shape[0] > max_degree:
            cut = np.sort(row)[max_degree]
            row[row > cut] = 0.0
    g.eliminate_zeros()
    g = (g != 0).astype(np.uint8)                                                      # :1613
    vertex_order = tree_idx                                                            # :1630
    g = g[vertex_order, :].tocsc()[:, vertex_order].tocsr(); g.sort_indices()          # :1631-1634
    data = np.ascontiguousarray(X[vertex_order, :])                                    # :1639-1641
    tree_order = np.argsort(vertex_order)
    new_tree_idx = tree_idx[tree_order].astype(np.int32)                               # :763-771 resort_tree_indices
    return dict(data=data, indptr=g.indptr.astype(np.int32), indices=g.indices.astype(np.int32), hyperplanes=hyper, offsets=offs,
                children=children, tree_indices=new_tree_idx, vertex_order=vertex_order.astype(np.int32),
                min_distance=np.float32(min_distance))


def correction(metric, d):
    """distances.py:2170-2188 fast_distance_alternatives[...]['correction'] (f64 output for the
    pow-based cosine/dot correction, as the reference's numba.vectorize gives)."""
    if metric in ("euclidean", "l2"):
        return np.sqrt(d)
    if metric in ("cosine", "dot", "jaccard"):            # correct_alternative_cosine / _jaccard: 1 - 2^-d
        return 1.0 - np.power(2.0, -d.astype(np.float64))
    if metric == "hellinger":                             # correct_alternative_hellinger: sqrt(1 - 2^-d)
        return np.sqrt(1.0 - np.power(2.0, -d.astype(np.float64)))
    if metric == "true_angular":                          # true_angular_from_alt_cosine (distances.py:1273-1280)
        return 1.0 - np.arccos(np.power(2.0, -d.astype(np.float64))) / np.pi
    if metric == "inner_product":                         # correct_alternative_inner_product (distances.py:841-853)
        d64 = d.astype(np.float64)
        return np.where(d >= np.finfo(np.float32).max, 0.0, -1.0 / np.where(d64 == 0.0, 1.0, d64))
    return d.copy()


class OracleNNDescent:
    """Replays NNDescent.__init__ (pynndescent_.py:976-1269) over the C oracle.

    n_threads is the numba thread count being emulated (it only matters inside
    new_build_candidates, utils.py:259-263).  `forest` = "dfs" (reference rng order) or
    "level" (node-hash rng, the GPU forest's scheme).  Intermediates are kept on `self`.
    """

    def __init__(self, data, metric="euclidean", n_neighbors=30, n_trees=None, leaf_size=None, random_state=None,
                 max_candidates=None, n_iters=None, delta=0.001, tree_init=True, max_rptree_depth=200, n_threads=8,
                 forest="dfs", leaf_array=None, keep_stages=False, block_size=16384, init_graph=None, init_dist=None,
                 metric_kwds=None):
        from sklearn.utils import check_random_state
        set_metric_param(metric, metric_kwds)
        data = np.ascontiguousarray(data, np.float32)
        n = data.shape[0]
        if n_trees is None:
            n_trees = max(3, min(12, int(round(2.0 * np.log10(n)))))
        if n_iters is None:
            n_iters = max(5, int(round(np.log2(n))))
        self.metric = metric
        self.n_neighbors = n_neighbors
        self.random_state = random_state
        self.n_trees = n_trees
        self.n_trees_after_update = max(2, int(np.round(n_trees / 3)))                 # pynndescent_.py:1016
        self.leaf_size, self.n_iters, self.delta, self.max_rptree_depth = leaf_size, n_iters, delta, max_rptree_depth
        self.n_threads, self.forest_scheme, self.block_size = n_threads, forest, block_size
        rs = check_random_state(random_state)
        if metric == "dot":
            from sklearn.preprocessing import normalize  # pynndescent_.py:1101-1103
            data = normalize(data, norm="l2", copy=True)
        self._raw_data = data
        self.rng_state = rs.randint(INT32_MIN, INT32_MAX, 3).astype(np.int64)
        self.search_rng_state = rs.randint(INT32_MIN, INT32_MAX, 3).astype(np.int64)
        angular = metric in ANGULAR_METRICS
        if leaf_array is not None:
            self.leaf_array = np.ascontiguousarray(leaf_array, np.int32)
        elif tree_init and n_trees > 0:
            self.leaf_array, self.tree_rng_states = make_forest(data, n_neighbors, n_trees, leaf_size, rs, angular,
                                                                max_rptree_depth, forest)
        else:
            self.leaf_array = np.array([[-1]], np.int32)
        C_ = min(60, n_neighbors) if max_candidates is None else max_candidates
        self.max_candidates = C_
        heap = make_heap(n, n_neighbors)
        if init_graph is not None:
            # pynndescent_.py:1225-1242 + nn_descent :339-350: the supplied graph REPLACES leaf init and init_random
            if np.shape(init_graph)[0] != n:
                raise ValueError("Init graph size does not match dataset size!")
            if init_dist is not None and np.shape(init_graph) != np.shape(init_dist):
                raise ValueError("The shapes of init graph and init distances do not match!")
            init_from_graph(data, metric, heap, init_graph, init_dist)
        elif tree_init:
            init_rp_tree(data, metric, heap, self.leaf_array)
        if keep_stages:
            self.heap_after_leaf = tuple(a.copy() for a in heap)
        if init_graph is None:
            init_random(data, metric, heap, self.rng_state)
        if keep_stages:
            self.heap_after_init = tuple(a.copy() for a in heap)
            self.rng_state_after_init = self.rng_state.copy()
        self.c_hist = np.zeros(n_iters, np.int64)
        self.n_iters_run = lib().orc_nn_descent_internal(data, n, data.shape[1], METRIC_IDS[metric], heap[0], heap[1],
                                                         heap[2], n_neighbors, C_, self.rng_state, n_threads, n_iters,
                                                         delta, block_size, self.c_hist)
        if keep_stages:
            self.heap_final = tuple(a.copy() for a in heap)
        deheap_sort(heap[0], heap[1])
        self._neighbor_graph = (heap[0], heap[1])

    def update(self, xs_fresh=None, xs_updated=None, updated_indices=None, leaf_array=None):
        """pynndescent_.py:2381-2553 update() (dense): replace rows, append rows, invalidate the graph entries that
        touch a replaced row, then forest (n_trees_after_update trees) -> init_from_neighbor_graph -> init_rp_tree ->
        nn_descent_internal on the index's own (already advanced) rng_state.  `leaf_array` injects a forest."""
        from sklearn.utils import check_random_state
        rs = check_random_state(self.random_state)                                      # :2407-2410
        rs.randint(INT32_MIN, INT32_MAX, 3)                                             # rng_state handed to make_forest (unused there)
        if xs_updated is not None and updated_indices is None:
            raise ValueError("If xs_updated are provided, updated_indices must also be provided!")
        upd = [] if (xs_updated is None or updated_indices is None) else list(map(int, updated_indices))
        data = self._raw_data.copy()
        if upd:
            xs_updated = np.ascontiguousarray(xs_updated, np.float32)
            if len(upd) != xs_updated.shape[0]:
                raise ValueError("Number of updated indices must match number of rows of xs_updated.")
            for x, i in zip(xs_updated, upd):                                           # :2472-2473
                data[i] = x
        if xs_fresh is not None and len(xs_fresh):
            data = np.vstack([data, np.ascontiguousarray(xs_fresh, np.float32)])        # :2474
        data = np.ascontiguousarray(data, np.float32)
        ns, ds = self._neighbor_graph[0].copy(), self._neighbor_graph[1].copy()
        if upd:                                                                         # :2475-2488
            changed = np.zeros(data.shape[0], bool); changed[upd] = True
            hit = changed[np.where(ns >= 0, ns, 0)] & (ns >= 0)
            hit[np.asarray(upd, np.int64)[np.asarray(upd) < ns.shape[0]], :] = True
            ns[hit] = -1; ds[hit] = np.inf
        self.n_trees = self.n_trees_after_update                                        # :2498
        angular = self.metric in ANGULAR_METRICS
        if leaf_array is not None:
            self.leaf_array = np.ascontiguousarray(leaf_array, np.int32)
        else:
            self.leaf_array, self.tree_rng_states = make_forest(data, self.n_neighbors, self.n_trees, self.leaf_size, rs,
                                                                angular, self.max_rptree_depth, self.forest_scheme)
        n = data.shape[0]
        heap = make_heap(n, self.n_neighbors)                                           # :2511
        init_from_neighbor_graph(heap, ns, ds)                                          # :2512-2514
        init_rp_tree(data, self.metric, heap, self.leaf_array)                          # :2515
        self.c_hist = np.zeros(self.n_iters, np.int64)
        self.n_iters_run = lib().orc_nn_descent_internal(data, n, data.shape[1], METRIC_IDS[self.metric], heap[0], heap[1],
                                                         heap[2], self.n_neighbors, self.max_candidates, self.rng_state,
                                                         self.n_threads, self.n_iters, self.delta, self.block_size,
                                                         self.c_hist)                   # :2522-2535
        deheap_sort(heap[0], heap[1])
        self._raw_data = data
        self._neighbor_graph = (heap[0], heap[1])

    @property
    def neighbor_graph(self):
        return self._neighbor_graph[0].copy(), correction(self.metric, self._neighbor_graph[1])
