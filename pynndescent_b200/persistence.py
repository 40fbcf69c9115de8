"""Index persistence in the REFERENCE's pickle layout.

`pynndescent.NNDescent.__getstate__` (pynndescent_.py:1306-1320) pickles the instance `__dict__` with the search forest
"de-numba-ified" to plain tuples (rp_trees.py:3060-3069); `__setstate__` (:1322-1331) restores it, re-resolves the metric
(`_set_distance_func`) and re-creates the numba search closure (`_init_search_function`).  `reference_state()` builds that
dict from a prepared GPU index and `dump_reference_pickle()` writes a pickle that a STOCK pynndescent installation loads
with `pickle.load` and queries on the CPU -- the end-to-end check that the graph / tree / ordering formats produced on
the GPU are the reference's own (tests/test_persistence.py runs exactly that against the real reference).

Host-side glue only: nothing here touches the device.
"""
import pickle
import sys
import types

import numpy as np

_REF_MODULE = "pynndescent.pynndescent_"


def _max_leaf_size(children):
    """FlatTree.leaf_size (rp_trees.py:2926-3049): leaves are rows with children[i, 0] <= 0 holding (-start, -end)."""
    c = np.asarray(children)
    leaf = c[:, 0] <= 0
    return int(np.max(c[leaf, 0] - c[leaf, 1])) if leaf.any() else 0


def reference_state(index):
    """The reference's `__getstate__` dict for a prepared `pynndescent_b200.NNDescent` (prepare() is run if needed)."""
    import scipy.sparse as sp
    if getattr(index, "_search_structures", None) is None:
        index.prepare()
    S = index._search_structures
    n = index._raw_data.shape[0]
    vertex_order = np.asarray(S["vertex_order"], np.int32)
    indptr = np.asarray(S["indptr"], np.int32); indices = np.asarray(S["indices"], np.int32)
    search_graph = sp.csr_matrix((np.ones(indices.shape[0], np.uint8), indices, indptr), shape=(n, n))   # :1613
    tree = (np.ascontiguousarray(S["hyperplanes"], np.float32), np.ascontiguousarray(S["offsets"], np.float32),
            np.ascontiguousarray(S["children"], np.int32), np.ascontiguousarray(S["tree_indices"], np.int32),
            _max_leaf_size(S["children"]))
    ng_idx = index._neighbor_graph[0]
    # the reference keeps the RAW (alternative-metric, uncorrected) distances in _neighbor_graph (pynndescent_.py:1247-1260;
    # the correction is applied by the `neighbor_graph` property): those are the sorted heap rows still on the device
    ng_dist = getattr(index, "_raw_graph_dist", None)
    if ng_dist is None:
        ng_dist = index._engine.heap_dist.cpu().numpy()
    state = dict(
        # constructor attributes (pynndescent_.py:1009-1052)
        n_trees=index.n_trees, n_trees_after_update=max(2, int(np.round(index.n_trees / 3))), n_neighbors=index.n_neighbors,
        metric=index.metric, metric_kwds=index.metric_kwds, bit_metric=False, angular_trees=False, leaf_size=index.leaf_size,
        prune_degree_multiplier=index.pruning_degree_multiplier, diversify_prob=1.0, diversify_method="standard",
        degree_prune_aggressiveness=1.0, n_search_trees=1, search_tree_leaf_size=getattr(index, "search_tree_leaf_size", None),
        max_search_tree_depth=getattr(index, "max_search_tree_depth", None),
        quantization=None, tree_init=index.tree_init, random_state=index.random_state, low_memory=index.low_memory,
        max_candidates=index.max_candidates, max_rptree_depth=index.max_rptree_depth, n_iters=index.n_iters, delta=index.delta,
        n_jobs=index.n_jobs, compressed=index.compressed, parallel_batch_queries=False, verbose=False, dim=index.dim,
        # derived state
        # :1064-1065, 1057, 1074-1090; _distance_correction is preset to None (:1296) -- the reference's __setstate__ /
        # _set_distance_func only overwrites it for metrics that have a fast alternative
        _dist_args=tuple((index.metric_kwds or {}).values()), _is_sparse=False,
        _input_dtype=np.dtype(np.float32), _angular_trees=bool(index._angular_trees), _distance_correction=None,
        _bit_trees=False, rng_state=np.asarray(index.rng_state, np.int64), search_rng_state=np.asarray(index.search_rng_state, np.int64),
        _raw_data=np.ascontiguousarray(index._raw_data[vertex_order, :]),                                    # :1639-1641
        _neighbor_graph=(np.asarray(ng_idx, np.int32), np.asarray(ng_dist, np.float32)),
        _search_graph=search_graph, _search_forest=(tree,), _vertex_order=vertex_order,
        _min_distance=np.float32(S["min_distance"]), _visited=np.zeros((n // 8) + 1, dtype=np.uint8, order="C"),  # :1657
    )
    return state


class _ReferenceClassStub:
    """Stands in for pynndescent.pynndescent_.NNDescent while PICKLING on a machine without pynndescent: pickle stores a
    class by (module, qualname) and only checks that the name resolves at dump time."""


def dump_reference_pickle(index, file, protocol=pickle.DEFAULT_PROTOCOL):
    """Write a pickle of `index` that `pickle.load` turns into a stock `pynndescent.NNDescent` (CPU) wherever the
    reference is installed.  `file` is a path or a binary file object."""
    state = reference_state(index)
    created = []
    try:
        try:
            import importlib
            cls = importlib.import_module(_REF_MODULE).NNDescent
        except Exception:
            for name in ("pynndescent", _REF_MODULE):
                if name not in sys.modules:
                    sys.modules[name] = types.ModuleType(name); created.append(name)
            stub = type("NNDescent", (_ReferenceClassStub,), {"__module__": _REF_MODULE, "__qualname__": "NNDescent"})
            sys.modules[_REF_MODULE].NNDescent = stub
            cls = stub
        payload = _Reduced(cls, state)
        if hasattr(file, "write"):
            pickle.dump(payload, file, protocol=protocol)
        else:
            with open(file, "wb") as f:
                pickle.dump(payload, f, protocol=protocol)
    finally:
        for name in created:
            sys.modules.pop(name, None)


class _Reduced:
    """Pickles as `object.__new__(cls)` followed by `__setstate__(state)` -- what unpickling a reference instance
    performs -- without needing an instance of the class (or the reference's own __getstate__) at dump time."""

    def __init__(self, cls, state):
        self.cls, self.state = cls, state

    def __reduce__(self):
        import copyreg
        return (copyreg._reconstructor, (self.cls, object, None), self.state)
