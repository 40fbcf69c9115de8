"""Seam 3 host side: the query graph walk over prepared search structures (pynndescent_.py:2275-2379 query()).

The search structures (pruned CSR search graph, flat hub/RP search tree, data re-ordered to leaf order) are the
product of the reference's `prepare()` (pynndescent_.py:1333-1662), which is a "next" row of the scope table;
this class takes them in the reference's own array format and runs `search_closure` + `deheap_sort` + the
`_vertex_order` / distance-correction epilogue on the GPU.
"""
import ctypes as C

import numpy as np
import torch

from . import distances as pynnd_dist
from ._ffi import SearchIndex, check, lib


def _dev(a, dtype, device):
    return torch.from_numpy(np.ascontiguousarray(a, dtype)).to(device)


class SearchIndexGPU:
    def __init__(self, data_reordered, metric, indptr, indices, hyperplanes, offsets, children, tree_indices, vertex_order,
                 min_distance, n_neighbors, search_rng_state, device=None, data_device=None, dim=None, metric_param=None):
        """`data_reordered`: host matrix in leaf order (pynndescent_.py:1639-1641) -- or, from prepare(), `data_device`: the same
        matrix already on the GPU as float32[N, ld] (rows padded to a multiple of 4 floats with zeros) with `dim` = d, so that
        the 4*N*d bytes neither get gathered on the host nor cross PCIe a second time."""
        if not torch.cuda.is_available():
            raise RuntimeError("pynndescent_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.metric = metric
        self.metric_id, self.dev_correction, self.host_correction, _ = pynnd_dist.resolve(metric)
        if data_device is not None:
            self.device = data_device.device
            self.N, self.ld = int(data_device.shape[0]), int(data_device.shape[1])
            self.d = int(dim)
            if self.ld != (self.d + 3) // 4 * 4 or data_device.dtype != torch.float32 or not data_device.is_contiguous():
                raise ValueError("data_device must be a contiguous float32[N, ld] tensor with ld = d rounded up to 4")
        else:
            self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
            data = np.ascontiguousarray(data_reordered, np.float32)
            self.N, self.d = data.shape
            self.ld = (self.d + 3) // 4 * 4
        self._ctx = C.c_void_p()
        check(lib.nnd_create(C.byref(self._ctx), self.device.index or 0))
        self.stream = torch.cuda.current_stream(self.device)
        if data_device is not None:
            self.X = data_device
        else:
            self.X = torch.zeros((self.N, self.ld), dtype=torch.float32, device=self.device)
            self.X[:, : self.d].copy_(torch.from_numpy(data))
        nb = C.c_size_t()
        check(lib.nnd_metric_aux_bytes(self.metric_id, self.N, C.byref(nb)))
        self.aux = torch.empty(nb.value, dtype=torch.uint8, device=self.device)
        check(lib.nnd_metric_prepare(self._ctx, C.c_void_p(self.X.data_ptr()), self.N, self.d, self.ld, self.metric_id,
                                     C.c_void_p(self.aux.data_ptr()), C.c_void_p(self.stream.cuda_stream)))
        if metric_param is not None:
            check(lib.nnd_metric_set_param(self._ctx, self.metric_id, C.c_void_p(self.aux.data_ptr()), float(metric_param),
                                           C.c_void_p(self.stream.cuda_stream)))
        hp = np.zeros((np.asarray(hyperplanes).shape[0], self.ld), np.float32)
        hp[:, : self.d] = np.asarray(hyperplanes, np.float32)[:, : self.d]
        self.indptr = _dev(indptr, np.int32, self.device); self.indices = _dev(indices, np.int32, self.device)
        self.hyper = _dev(hp, np.float32, self.device); self.offsets = _dev(offsets, np.float32, self.device)
        self.children = _dev(children, np.int32, self.device); self.tree_indices = _dev(tree_indices, np.int32, self.device)
        self.vertex_order = _dev(vertex_order, np.int64, self.device)
        self.min_distance = float(min_distance)
        self.n_neighbors = int(n_neighbors)
        self.search_rng_state = np.asarray(search_rng_state, np.int64).copy()
        self.last_stats = None

    def close(self):
        if self._ctx:
            lib.nnd_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def query(self, query_data, k=10, epsilon=0.1):
        """pynndescent_.py:2275-2379 (dense, non-proxy): returns (indices in input order, corrected distances)."""
        Q = np.asarray(query_data).astype(np.float32, order="C")                       # :2319
        nq = Q.shape[0]
        Qd = torch.zeros((nq, self.ld), dtype=torch.float32, device=self.device)
        Qd[:, : self.d].copy_(torch.from_numpy(Q))
        out_i = torch.empty((nq, k), dtype=torch.int32, device=self.device)
        out_d = torch.empty((nq, k), dtype=torch.float32, device=self.device)
        nb = C.c_size_t()
        check(lib.nnd_query_workspace_bytes(self.N, nq, int(k), C.byref(nb)))
        ws = torch.empty(nb.value, dtype=torch.uint8, device=self.device)
        S = SearchIndex(self.X.data_ptr(), self.N, self.d, self.ld, self.metric_id, self.aux.data_ptr(), self.indptr.data_ptr(),
                        self.indices.data_ptr(), self.hyper.data_ptr(), self.offsets.data_ptr(), self.children.data_ptr(),
                        self.tree_indices.data_ptr(), self.min_distance, self.n_neighbors, int(self.metric in ("cosine", "dot")))
        rng = (C.c_int64 * 3)(*[int(v) for v in self.search_rng_state])              # :1796 copy per call
        stats = (C.c_int64 * 3)()
        st = C.c_void_p(self.stream.cuda_stream)
        check(lib.nnd_query(self._ctx, C.byref(S), C.c_void_p(Qd.data_ptr()), nq, int(k), float(epsilon), rng,
                            C.c_void_p(out_i.data_ptr()), C.c_void_p(out_d.data_ptr()), C.c_void_p(ws.data_ptr()), nb.value,
                            stats, st))
        self.last_stats = {"distance_evals": int(stats[0]), "pops": int(stats[1]), "sequential_queries": int(stats[2])}
        if nq >= 64 and self.last_stats["sequential_queries"] > 0.05 * nq:
            # queries that overflow the on-chip state (visited table / seed heap) or need the shared rng are replayed one after the
            # other by ONE warp (k_query_seq): correct, but throughput collapses when many of a batch take that path (large k / epsilon)
            from warnings import warn
            warn("%d of %d queries were replayed sequentially on the GPU (on-chip search state overflow or rng draws); expect low "
                 "throughput for this k / epsilon" % (self.last_stats["sequential_queries"], nq))
        sorted_d = torch.empty_like(out_d)
        check(lib.nnd_deheap_sort(self._ctx, C.c_void_p(out_i.data_ptr()), C.c_void_p(out_d.data_ptr()), nq, int(k),
                                  int(self.dev_correction), C.c_void_p(sorted_d.data_ptr()), st))           # :2361
        idx = self.vertex_order[out_i.long()].to(torch.int32)                                             # :2374
        self.stream.synchronize()
        dists = sorted_d.cpu().numpy()
        if self.host_correction is not None:                                                            # :2376-2377
            dists = self.host_correction(dists)
        return idx.cpu().numpy(), dists
