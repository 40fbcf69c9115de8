"""NNDescent -- host-side mirror of the reference's index object for the build path.

Same constructor keywords, same `.neighbor_graph` contract as
/root/reference/pynndescent/pynndescent_.py:792-1304, 2145-2158; the three numba seams
(make_forest, nn_descent, deheap_sort + correction) are replaced by libnnd_b200.so kernels.

Determinism contract: for a given (data, random_state) the result is a pure function of the
emulated numba thread count T (`n_jobs`, default 8 -- the reference's own result depends on its
thread count in exactly the same way, utils.py:259-263) and of nothing else: not of the GPU
count, block sizes or kernel scheduling.
"""
import time
from warnings import warn

import numpy as np

from . import distances as pynnd_dist
from .engine import Engine

INT32_MIN = np.iinfo(np.int32).min + 1
INT32_MAX = np.iinfo(np.int32).max - 1
DEFAULT_EMULATED_THREADS = 8


def _tau_advance(state, n):
    """utils.py:17-40 tau_rand_int applied n times (int64 state words, arithmetic >>)."""
    s = [int(v) for v in state]

    def wrap(x):
        x &= (1 << 64) - 1
        return x - (1 << 64) if x >= (1 << 63) else x
    for _ in range(n):
        s[0] = wrap((((s[0] & 4294967294) << 12) & 0xFFFFFFFF) ^ ((((s[0] << 13) & 0xFFFFFFFF) ^ s[0]) >> 19))
        s[1] = wrap((((s[1] & 4294967288) << 4) & 0xFFFFFFFF) ^ ((((s[1] << 2) & 0xFFFFFFFF) ^ s[1]) >> 25))
        s[2] = wrap((((s[2] & 4294967280) << 17) & 0xFFFFFFFF) ^ ((((s[2] << 3) & 0xFFFFFFFF) ^ s[2]) >> 11))
    return np.array(s, np.int64)


def ts():
    return time.ctime(time.time())


class NNDescent:
    """Approximate k-nearest-neighbour graph construction by NN-descent on one B200.

    Parameters mirror pynndescent.NNDescent (pynndescent_.py:976-1007).  Extras (optional):
    ``device`` (torch device string) and ``push_capacity`` (scratch sizing).  Keywords that select
    code outside the accelerated hot path (sparse input, callable metrics, quantisation, ...)
    raise NotImplementedError instead of silently running on the CPU.
    """

    def __init__(self, data, metric="euclidean", metric_kwds=None, bit_metric=False, n_neighbors=30, n_trees=None,
                 angular_trees=False, leaf_size=None, pruning_degree_multiplier=1.5, diversify_prob=1.0,
                 diversify_method="standard", degree_prune_aggressiveness=1.0, n_search_trees=1,
                 search_tree_leaf_size=None, max_search_tree_depth=None, quantization=None, tree_init=True,
                 init_graph=None, init_dist=None, random_state=None, low_memory=True, max_candidates=None,
                 max_rptree_depth=200, n_iters=None, delta=0.001, n_jobs=None, compressed=False,
                 parallel_batch_queries=False, verbose=False, device=None, push_capacity=None, leaf_array=None,
                 sharded=False):
        from sklearn.utils import check_random_state, check_array
        from scipy.sparse import issparse

        if issparse(data):
            raise NotImplementedError("sparse input is outside the sm_100a hot path (SURVEY.md 2 #13)")
        if quantization is not None or bit_metric:
            raise NotImplementedError("quantized / bit metrics are outside the sm_100a hot path")
        if init_graph is not None:                                                   # pynndescent_.py:1228-1236
            if np.shape(init_graph)[0] != data.shape[0]:
                raise ValueError("Init graph size does not match dataset size!")
            if init_dist is not None and np.shape(init_graph) != np.shape(init_dist):
                raise ValueError("The shapes of init graph and init distances do not match!")
            if sharded:
                raise NotImplementedError("init_graph with the sharded build")
        # keywords of prepare() whose non-default values select code that is not built here (pynndescent_.py:1462-1587):
        # never a silent drop.  `angular_trees` only matters for callable metrics in the reference (:1091-1095), which
        # raise here anyway; `parallel_batch_queries` / `degree_prune_aggressiveness` (degree_aware only) change no result.
        if diversify_prob != 1.0:
            raise NotImplementedError("diversify_prob != 1.0 (randomised diversification) is outside the sm_100a hot path")
        if diversify_method != "standard":
            raise NotImplementedError("diversify_method=%r (only 'standard' is built)" % (diversify_method,))
        if n_search_trees != 1:
            raise NotImplementedError("n_search_trees != 1")
        if leaf_array is not None and sharded:
            raise NotImplementedError("leaf_array together with the sharded build")
        if n_trees is None:
            n_trees = max(3, min(12, int(round(2.0 * np.log10(data.shape[0])))))      # pynndescent_.py:1009-1010
        if n_iters is None:
            n_iters = max(5, int(round(np.log2(data.shape[0]))))                     # pynndescent_.py:1011-1012
        self.n_trees = n_trees
        self.n_neighbors = n_neighbors
        self.metric = metric
        self.metric_kwds = metric_kwds
        self.leaf_size = leaf_size
        self.pruning_degree_multiplier = pruning_degree_multiplier
        self.angular_trees = angular_trees
        self.diversify_prob = diversify_prob
        self.diversify_method = diversify_method
        self.degree_prune_aggressiveness = degree_prune_aggressiveness
        self.n_search_trees = n_search_trees
        self.search_tree_leaf_size = search_tree_leaf_size
        self.max_search_tree_depth = max_search_tree_depth
        self.parallel_batch_queries = parallel_batch_queries
        self.max_candidates = max_candidates
        self.max_rptree_depth = max_rptree_depth
        self.n_iters = n_iters
        self.delta = delta
        self.n_jobs = n_jobs
        self.low_memory = low_memory
        self.compressed = compressed
        self.verbose = verbose
        self.random_state = random_state
        self.dim = data.shape[1]

        metric_id, dev_correction, host_correction, angular = pynnd_dist.resolve(metric)    # :1271-1292
        metric_param = pynnd_dist.metric_param(metric, metric_kwds)                          # :1064-1065, 1294-1302
        self._distance_correction = host_correction
        self._angular_trees = angular
        # :1054 check_array -- dtype / layout conversion on the host; the NaN/inf scan (sklearn sums the whole
        # matrix on one core, ~50 ms per 512 MB) is done on the device right after the H2D copy instead
        try:
            data = check_array(data, dtype=np.float32, order="C", ensure_all_finite=False)
        except TypeError:  # older scikit-learn
            data = check_array(data, dtype=np.float32, order="C", force_all_finite=False)
        self.tree_init = bool(tree_init) and n_trees != 0                                    # :1059-1062
        current_random_state = check_random_state(self.random_state)
        if metric == "dot":                                                                  # :1101-1103
            from sklearn.preprocessing import normalize
            data = normalize(data, norm="l2", copy=True)
        self._raw_data = data
        self.rng_state = current_random_state.randint(INT32_MIN, INT32_MAX, 3).astype(np.int64)          # :1105-1107
        self.search_rng_state = current_random_state.randint(INT32_MIN, INT32_MAX, 3).astype(np.int64)  # :1108-1110
        # :1111-1113 warm up the search rng by 10 draws (host-side replica of tau_rand_int)
        self.search_rng_state = _tau_advance(self.search_rng_state, 10)
        T = DEFAULT_EMULATED_THREADS if (n_jobs is None or n_jobs == -1) else int(n_jobs)
        if T < 1:
            raise ValueError("n_jobs must be a positive thread count (or None / -1)")
        self.emulated_threads = T
        eff_C = min(60, n_neighbors) if max_candidates is None else max_candidates           # :1135-1138
        self.n_trees_after_update = max(2, int(np.round(self.n_trees / 3)))                  # :1016
        self._build_args = dict(metric_id=metric_id, metric_param=metric_param, dev_correction=dev_correction, eff_C=eff_C, device=device,
                                push_capacity=push_capacity, sharded=bool(sharded))

        t0 = time.perf_counter()
        eng = Engine(data, metric_id, n_neighbors, device=device, metric_param=metric_param, distributed_upload=bool(sharded))
        self._engine = eng
        if not eng.all_finite():
            raise ValueError("Input contains NaN, infinity or a value too large for dtype('float32').")
        self.timings = {"h2d_setup": time.perf_counter() - t0}
        tree_states = None
        if leaf_array is None and self.tree_init:
            if verbose:
                print(ts(), "Building RP forest with", str(n_trees), "trees")
            if leaf_size is None:
                leaf_size = max(60, min(256, 5 * int(n_neighbors)))                          # rp_trees.py:2845-2846
            tree_states = current_random_state.randint(INT32_MIN, INT32_MAX, size=(n_trees, 3)).astype(np.int64)
            self._tree_rng_states = tree_states                                              # rp_trees.py:2850-2852
        if verbose:
            print(ts(), "NN descent for", str(n_iters), "iterations")
        if sharded:
            # one process per GPU (torch.distributed, NCCL): rows of the heap are sharded, see sharded.py
            from .sharded import ShardedBuilder
            sb = ShardedBuilder(eng)
            self._sharded = sb
            idx_t, dist_t, iters, c_hist = sb.build(self.rng_state, tree_states, leaf_size, angular, eff_C, T, n_iters, delta,
                                                    dev_correction, max_depth=max_rptree_depth, cap=push_capacity,
                                                    timings=self.timings)
            stats = [0, 0, 0, 0]
        else:
            idx_t, dist_t, iters, c_hist, stats = eng.build(
                self.rng_state, tree_states, leaf_size, angular, eff_C, T, n_iters, delta, dev_correction,
                max_depth=max_rptree_depth, push_capacity=push_capacity, leaf_array=leaf_array, timings=self.timings,
                init_graph=init_graph, init_dist=init_dist if init_graph is not None else None)
        if verbose:
            for it in range(iters):
                print("\t", it + 1, " / ", n_iters)
            if c_hist and c_hist[-1] <= delta * n_neighbors * data.shape[0]:
                print("\tStopping threshold met -- exiting after", iters, "iterations")
        self.n_iters_run, self.c_hist, self.join_stats = iters, c_hist, stats
        self._set_graph(idx_t, dist_t)

    # ---- the kNN graph stays on the device; host copies are made when somebody asks for them
    def _set_graph(self, idx_t, dist_t):
        self.__dict__.pop("_neighbor_graph", None)
        self._graph_dev = (idx_t, dist_t)
        if bool((idx_t < 0).any()):                                                          # :1262-1267
            warn("Failed to correctly find n_neighbors for some samples."
                 " Results may be less than ideal. Try re-running with"
                 " different parameters.")

    def __getattr__(self, name):
        # `_neighbor_graph` (the reference's attribute: host (indices, distances), used by update() and persistence) is
        # fetched from the device on first use and cached; an explicitly assigned value shadows this
        if name == "_neighbor_graph":
            dev = self.__dict__.get("_graph_dev")
            if dev is not None:
                g = self.__dict__["_engine"].to_host(*dev)
                self.__dict__["_neighbor_graph"] = g
                return g
        raise AttributeError(f"{type(self).__name__!r} object has no attribute {name!r}")

    @property
    def neighbor_graph(self):
        """pynndescent_.py:2145-2158: (indices copy, corrected distances).  Every access returns arrays the caller owns:
        a fresh device -> host copy (one PCIe transfer into pinned staging, no second host-side memcpy), or copies of the
        cached host graph when one exists."""
        cached = self.__dict__.get("_neighbor_graph")
        if cached is not None:
            hi, hd, own = cached[0].copy(), cached[1], False
        elif self.__dict__.get("_graph_dev") is not None:
            (hi, hd), own = self._engine.to_host(*self._graph_dev), True
        else:
            raise AttributeError("'NNDescent' object has no attribute '_neighbor_graph'")      # compress_index() dropped it
        if self._distance_correction is not None:
            return hi, self._distance_correction(hd)
        return hi, (hd if own else hd.copy())

    def update(self, xs_fresh=None, xs_updated=None, updated_indices=None, leaf_array=None):
        """pynndescent_.py:2381-2553 (dense): append `xs_fresh`, replace rows `updated_indices` by `xs_updated`, then
        re-run the hot path from the old graph: every entry that touches a replaced row is invalidated (:2475-2488), a
        forest of n_trees_after_update trees is grown on the new data (:2498-2509), the heap is seeded with the old
        graph (init_from_neighbor_graph :207-214, unflagged pushes), refined by the leaf join (init_rp_tree) and by
        nn_descent on the index's own rng_state (:2522-2535).  All of it runs on the kernels of the build path.

        The old entries are re-pushed with their RAW distances.  The reference keeps those on the host; here they are
        recomputed on the device from the (index, index) pairs -- the canonical summation order makes that bit-identical
        to what the heap held -- because the host copy holds the corrected distances.  `leaf_array` (extra, optional)
        injects a forest in the reference's format, as in the constructor.
        """
        from sklearn.utils import check_array, check_random_state
        if self._build_args["sharded"]:
            raise NotImplementedError("update() of a sharded index")
        import inspect
        _nofinite = ({"ensure_all_finite": False} if "ensure_all_finite" in inspect.signature(check_array).parameters
                     else {"force_all_finite": False})
        current_random_state = check_random_state(self.random_state)                             # :2407
        current_random_state.randint(INT32_MIN, INT32_MAX, 3)     # :2408-2410 the rng_state make_forest receives and ignores
        if xs_updated is not None:                                                                 # :2413-2434
            xs_updated = check_array(xs_updated, dtype=np.float32, order="C", **_nofinite)
            if updated_indices is None:
                raise ValueError("If xs_updated are provided, updated_indices must also be provided!")
            try:
                updated_indices = list(map(int, updated_indices))
            except (TypeError, ValueError):
                raise ValueError("Could not convert updated indices to list of int(s).")
            if len(updated_indices) != xs_updated.shape[0]:
                raise ValueError(f"Number of updated indices ({len(updated_indices)}) must match "
                                 f"number of rows of xs_updated ({xs_updated.shape[0]}).")
        else:
            if updated_indices is not None:                                                        # :2435-2441
                warn("xs_updated not provided, while update_indices provided. They will be ignored.")
            updated_indices = None
        if updated_indices is None:
            xs_updated, updated_indices = [], []
        if xs_fresh is None:
            xs_fresh = np.zeros((0, self._raw_data.shape[1]), dtype=np.float32)                   # :2453-2456
        else:
            xs_fresh = check_array(xs_fresh, dtype=np.float32, order="C", **_nofinite)
        # check_array's NaN/inf rejection (:2414, :2458), on the rows that are new, BEFORE any state is touched
        if not (np.isfinite(xs_fresh).all() and (len(updated_indices) == 0 or np.isfinite(xs_updated).all())):
            raise ValueError("Input contains NaN, infinity or a value too large for dtype('float32').")
        data = np.array(self._raw_data, dtype=np.float32, order="C", copy=True)                   # :2471-2474
        for x_updated, i_fresh in zip(xs_updated, updated_indices):
            data[i_fresh] = x_updated
        data = np.ascontiguousarray(np.vstack([data, xs_fresh]))
        ns = self._neighbor_graph[0].copy()                                                       # :2475-2488
        if updated_indices:
            changed = np.zeros(data.shape[0], dtype=bool)
            changed[updated_indices] = True
            ns[(ns >= 0) & changed[np.where(ns >= 0, ns, 0)]] = -1
            ns[np.asarray(updated_indices, dtype=np.int64), :] = -1
        self.n_trees = self.n_trees_after_update                                                  # :2498
        a = self._build_args
        old = self._engine
        self._engine = None
        old.close()
        del old
        t0 = time.perf_counter()
        eng = Engine(data, a["metric_id"], self.n_neighbors, device=a["device"], metric_param=a["metric_param"])
        self._engine = eng
        self.timings = {"h2d_setup": time.perf_counter() - t0}
        cap = a["push_capacity"] or eng.default_push_capacity()
        eng.workspace(a["eff_C"], self.emulated_threads, cap)
        # :2511-2514 make_heap + init_from_neighbor_graph: the old rows in column order; -1 entries push nothing
        eng.init_from_graph(ns, None, cap)
        eng.heap_flag.zero_()                  # init_from_neighbor_graph pushes with flag 0 (the warm-start replay sets 1)
        if leaf_array is not None:                                                                # :2515 init_rp_tree
            eng.leaf_init_from_array(leaf_array, cap)
        else:
            leaf_size = self.leaf_size if self.leaf_size is not None else max(60, min(256, 5 * int(self.n_neighbors)))
            tree_states = current_random_state.randint(INT32_MIN, INT32_MAX, size=(self.n_trees, 3)).astype(np.int64)
            lp, lo, n_leaves, max_leaf = eng.forest(tree_states, leaf_size, self._angular_trees, self.max_rptree_depth)
            eng.leaf_init(lp, lo, n_leaves, max_leaf, cap)
            del lp, lo
        iters, c_hist, stats = eng.descent(a["eff_C"], self.rng_state, self.emulated_threads, self.n_iters, self.delta, cap)
        idx_t, dist_t = eng.deheap_sort(a["dev_correction"])
        self.n_iters_run, self.c_hist, self.join_stats = iters, c_hist, stats
        self._raw_data = data
        self._set_graph(idx_t, dist_t)
        self.timings["update"] = time.perf_counter() - t0
        if getattr(self, "_search_gpu", None) is not None:                                        # :2539-2553
            self._search_gpu = None
            self._search_structures = None
            self.prepare()

    def compress_index(self):
        """pynndescent_.py:2160-2172: prepare the search structures, then drop what only construction needed (the host
        copy of the kNN graph; `neighbor_graph` raises AttributeError afterwards exactly like the reference's)."""
        import gc
        self.prepare()
        self.compressed = True
        self.__dict__.pop("_neighbor_graph", None)
        self.__dict__.pop("_graph_dev", None)
        gc.collect()

    def prepare(self):
        """pynndescent_.py:2174-2273 (dense, no quantisation): hub search tree + diversified, degree-pruned search graph,
        data re-ordered to leaf order; then the GPU search index (seam 3)."""
        if getattr(self, "_search_gpu", None) is not None:
            return
        if self.pruning_degree_multiplier is None:
            raise ValueError("pruning_degree_multiplier must be a number")
        if not self.tree_init:
            # the reference then searches WITHOUT a tree (_search_forest = [], pynndescent_.py:1373-1374): not built here
            raise NotImplementedError("prepare() / query() of an index built with tree_init=False (tree-less search)")
        from .prepare import prepare_search_structures
        eng = self._engine
        # :1341-1351 dedicated search-tree parameters win over the construction ones
        search_leaf_size = self.search_tree_leaf_size if self.search_tree_leaf_size is not None else self.leaf_size
        search_depth = self.max_search_tree_depth if self.max_search_tree_depth is not None else self.max_rptree_depth
        S = prepare_search_structures(eng, self.metric, eng.heap_idx, eng.heap_dist, self.n_neighbors, self._angular_trees,
                                      leaf_size=search_leaf_size, pruning_degree_multiplier=self.pruning_degree_multiplier,
                                      max_depth=search_depth)
        self._search_structures = S
        # :1639-1641 data re-ordered to leaf order -- gathered on the device from the matrix that is already there
        import torch
        from .search import SearchIndexGPU
        vo_t = torch.from_numpy(np.ascontiguousarray(S["vertex_order"], np.int64)).to(eng.device)
        self._search_gpu = SearchIndexGPU(None, self.metric, S["indptr"], S["indices"], S["hyperplanes"], S["offsets"], S["children"],
                                          S["tree_indices"], S["vertex_order"], S["min_distance"], self.n_neighbors,
                                          self.search_rng_state, data_device=eng.X.index_select(0, vo_t).contiguous(), dim=eng.d,
                                          metric_param=self._build_args["metric_param"])

    def load_search_structures(self, data_reordered, indptr, indices, hyperplanes, offsets, children, tree_indices,
                               vertex_order, min_distance, search_rng_state=None):
        """Install search structures in the reference's array format (what prepare() computes,
        pynndescent_.py:1670-1673, 1746-1750, 1629-1641, 1539) so that query() can run on the GPU."""
        from .search import SearchIndexGPU
        self._search_gpu = SearchIndexGPU(data_reordered, self.metric, indptr, indices, hyperplanes, offsets, children,
                                          tree_indices, vertex_order, min_distance, self.n_neighbors,
                                          self.search_rng_state if search_rng_state is None else search_rng_state,
                                          metric_param=pynnd_dist.metric_param(self.metric, self.metric_kwds))
        return self

    def reference_state(self):
        """The reference's pickle-state dict (pynndescent_.py:1306-1320) for this index; see persistence.py."""
        from .persistence import reference_state
        return reference_state(self)

    def save_for_reference(self, file):
        """Pickle this index so that a stock pynndescent installation can `pickle.load` and query it on the CPU."""
        from .persistence import dump_reference_pickle
        dump_reference_pickle(self, file)

    def query(self, query_data, k=10, epsilon=0.1):
        """pynndescent_.py:2275-2379."""
        self.prepare()
        return self._search_gpu.query(query_data, k=k, epsilon=epsilon)
