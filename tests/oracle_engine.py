"""TEST INFRASTRUCTURE: an Engine look-alike whose stages run on the CPU oracle instead of libnnd_b200's kernels.

It exists so that the HOST logic of pynndescent_b200.NNDescent (constructor, update()) can be exercised in the CPU suite
(`-m "not gpu"`): the tests monkeypatch `pynndescent_b200.pynndescent_.Engine` with this class.  It implements exactly the
subset of pynndescent_b200.engine.Engine those two code paths call, with the same argument meaning and the same
in-place side effects (rng_state advanced by init_random, heaps owned by the engine).  Never imported by the product.
"""
import numpy as np
import torch

from oracle import nnd_oracle as orc

ID_TO_NAME = {0: "sqeuclidean", 1: "cosine", 2: "dot", 3: "correlation", 4: "manhattan", 5: "chebyshev", 6: "canberra",
              7: "hamming", 8: "jaccard", 16: "hellinger", 17: "inner_product", 18: "minkowski"}


class OracleEngine:
    def __init__(self, data, metric_id, n_neighbors, device=None, metric_param=None, distributed_upload=False):
        if metric_param is not None:
            orc.lib().orc_set_metric_param(float(metric_param))
        self.X = np.ascontiguousarray(data, np.float32)
        self.N, self.d = self.X.shape
        self.k = int(n_neighbors)
        self.metric = ID_TO_NAME[int(metric_id)]
        self.closed = False
        self.calls = []
        self.make_heap()

    # ---- plumbing
    def close(self):
        self.closed = True

    def all_finite(self):
        return bool(np.isfinite(self.X).all())

    def default_push_capacity(self):
        return 1 << 23

    def workspace(self, C_, T, cap):
        return None

    def launch_count(self):
        return 0

    def to_host(self, idx_t, dist_t):
        return idx_t.numpy().copy(), dist_t.numpy().copy()

    @property
    def heap(self):
        return (self.heap_idx.numpy(), self.heap_dist.numpy(), self.heap_flag.numpy())

    def pair_distances(self, a, b):
        return np.array([orc.dist(self.metric, self.X[i], self.X[j]) for i, j in zip(a, b)], np.float32)

    def get_heap(self):
        return tuple(a.copy() for a in self.heap)

    def set_heap(self, idx, dist, flag):
        hi, hd, hf = self.heap
        hi[...] = idx; hd[...] = dist; hf[...] = flag

    def build_candidates(self, max_candidates, rng_state, n_threads, tables_valid=False):
        # (the engine's rng_state argument is advanced in place, like nnd_build_candidates')
        newc, oldc = orc.new_build_candidates(self.heap, int(max_candidates), rng_state, int(n_threads))
        return torch.from_numpy(newc), torch.from_numpy(oldc)

    def process_candidates(self, newc, oldc, push_capacity=None):
        c, em = orc.process_candidates(self.X, self.metric, self.heap, np.ascontiguousarray(newc.numpy()), np.ascontiguousarray(oldc.numpy()), 8)
        return c, [0, 0, em, 0]

    # ---- stages
    def make_heap(self):
        hi, hd, hf = orc.make_heap(self.N, self.k)
        self.heap_idx, self.heap_dist, self.heap_flag = torch.from_numpy(hi), torch.from_numpy(hd), torch.from_numpy(hf)

    def init_from_graph(self, init_graph, init_dist=None, push_capacity=None):
        self.calls.append("init_from_graph")
        g = np.ascontiguousarray(init_graph, np.int32)
        if g.size and int(g.max()) >= self.N:
            raise ValueError("init_graph refers to a point outside the dataset")
        # Engine.init_from_graph accepts fewer rows than the heap (update(): the old graph); rows beyond stay empty
        full = np.full((self.N, g.shape[1]), -1, np.int32); full[: g.shape[0]] = g
        dd = None
        if init_dist is not None:
            dd = np.full((self.N, g.shape[1]), np.inf, np.float32); dd[: g.shape[0]] = init_dist
        orc.init_from_graph(self.X, self.metric, self.heap, full, dd)

    def leaf_init_from_array(self, leaf_array, push_capacity=None):
        self.calls.append("leaf_init_from_array")
        orc.init_rp_tree(self.X, self.metric, self.heap, np.ascontiguousarray(leaf_array, np.int32))

    def init_random(self, rng_state):
        orc.init_random(self.X, self.metric, self.heap, rng_state)

    def descent(self, max_candidates, rng_state, n_threads, n_iters, delta, push_capacity=None):
        self.calls.append("descent")
        hi, hd, hf = self.heap
        c_hist = np.zeros(n_iters, np.int64)
        it = orc.lib().orc_nn_descent_internal(self.X, self.N, self.d, orc.METRIC_IDS[self.metric], hi, hd, hf, self.k,
                                               int(max_candidates), np.ascontiguousarray(rng_state, np.int64), int(n_threads),
                                               int(n_iters), float(delta), 16384, c_hist)
        return int(it), [int(c) for c in c_hist[:it]], [0, 0, 0, 0]

    def deheap_sort(self, correction):
        hi, hd, _ = self.heap
        orc.deheap_sort(hi, hd)
        out = np.sqrt(hd) if int(correction) == 1 else hd.copy()
        return self.heap_idx, torch.from_numpy(out)

    def build(self, rng_state, tree_rng_states, leaf_size, angular, max_candidates, n_threads, n_iters, delta, correction,
              max_depth=200, push_capacity=None, leaf_array=None, timings=None, init_graph=None, init_dist=None):
        self.make_heap()
        if init_graph is not None:
            self.init_from_graph(init_graph, init_dist)
        else:
            if leaf_array is None and tree_rng_states is not None and len(tree_rng_states):
                leaf_array = orc.forest_leaf_array(self.X, np.ascontiguousarray(tree_rng_states, np.int64), leaf_size, angular,
                                                   max_depth, "level")
            if leaf_array is not None:
                self.leaf_init_from_array(leaf_array)
            self.init_random(rng_state)
        iters, c_hist, stats = self.descent(max_candidates, rng_state, n_threads, n_iters, delta)
        idx_t, dist_t = self.deheap_sort(correction)
        return idx_t, dist_t, iters, c_hist, stats
