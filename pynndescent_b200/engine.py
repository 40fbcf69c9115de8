"""Stage-level driver over the C-ABI (include/nnd_b200.h).

PyTorch is used ONLY as the device-array container (torch.empty / data_ptr / streams); every
computation is a kernel of libnnd_b200.so.  The stage methods map 1:1 onto the numba functions
of the reference so that tests can inject the reference's intermediates at any seam:

    forest()             rp_trees.py:2815-2922    make_forest + rptree_leaf_array
    leaf_init()          pynndescent_.py:116-185  init_rp_tree
    init_random()        pynndescent_.py:188-203
    build_candidates()   utils.py:221-320         new_build_candidates
    process_candidates() pynndescent_.py:217-263
    descent()            pynndescent_.py:266-320  nn_descent_internal
    deheap_sort()        utils.py:189-218
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _ffi
from ._ffi import Problem, check, lib


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Engine:
    def __init__(self, data, metric_id, n_neighbors, device=None, metric_param=None, distributed_upload=False):
        if not torch.cuda.is_available():
            raise RuntimeError("pynndescent_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        torch.cuda.set_device(self.device)
        self._ctx = C.c_void_p()
        check(lib.nnd_create(C.byref(self._ctx), self.device.index or 0))
        data = np.ascontiguousarray(data, dtype=np.float32)
        self.N, self.d = data.shape
        self.ld = (self.d + 3) // 4 * 4
        self.k = int(n_neighbors)
        self.metric_id = int(metric_id)
        self.stream = torch.cuda.current_stream(self.device)
        host = torch.from_numpy(data)
        world = rank = None
        if distributed_upload:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                world, rank = dist.get_world_size(), dist.get_rank()
        if world is None:
            if self.ld == self.d:
                self.X = torch.empty((self.N, self.ld), dtype=torch.float32, device=self.device)
                self.X.copy_(host, non_blocking=host.is_pinned())
            else:  # pad rows to a multiple of 4 floats (16-byte vector loads / bulk copies); pad columns are zero
                self.X = torch.zeros((self.N, self.ld), dtype=torch.float32, device=self.device)
                self.X[:, : self.d].copy_(host, non_blocking=host.is_pinned())
        else:
            # sharded build: X is replicated on every GPU, but every process holds the same host matrix -- each uploads 1 / world of
            # the rows over PCIe and the GPUs complete their copies over NVLink (one all-gather) instead of world full uploads
            import torch.distributed as dist
            rows = (self.N + world - 1) // world
            buf = torch.zeros((world * rows, self.ld), dtype=torch.float32, device=self.device)
            lo, hi = min(rank * rows, self.N), min((rank + 1) * rows, self.N)
            if hi > lo:
                buf[lo:hi, : self.d].copy_(host[lo:hi], non_blocking=host.is_pinned())
            dist.all_gather_into_tensor(buf, buf[rank * rows: (rank + 1) * rows].clone())
            self.X = buf[: self.N]
        nb = C.c_size_t()
        check(lib.nnd_metric_aux_bytes(self.metric_id, self.N, C.byref(nb)))
        self.aux = torch.empty(nb.value, dtype=torch.uint8, device=self.device)
        check(lib.nnd_metric_prepare(self._ctx, _ptr(self.X), self.N, self.d, self.ld, self.metric_id, _ptr(self.aux),
                                     self._st()))
        if metric_param is not None:      # distances.py:138 minkowski's p (metric_kwds.values(), pynndescent_.py:1294-1302)
            check(lib.nnd_metric_set_param(self._ctx, self.metric_id, _ptr(self.aux), float(metric_param), self._st()))
        self.heap_idx = torch.empty((self.N, self.k), dtype=torch.int32, device=self.device)
        self.heap_dist = torch.empty((self.N, self.k), dtype=torch.float32, device=self.device)
        self.heap_flag = torch.empty((self.N, self.k), dtype=torch.uint8, device=self.device)
        self._ws = None
        self._ws_key = None
        self.make_heap()

    # ------------------------------------------------------------------ plumbing
    def _st(self):
        return C.c_void_p(self.stream.cuda_stream)

    def close(self):
        if self._ctx:
            lib.nnd_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def problem(self):
        return Problem(self.X.data_ptr(), self.N, self.d, self.ld, self.metric_id, self.aux.data_ptr(), self.k,
                       self.heap_idx.data_ptr(), self.heap_dist.data_ptr(), self.heap_flag.data_ptr())

    def default_push_capacity(self):
        # 56 B of routing workspace per record; 2^29 records (30 GB) keep a 10M-point iteration to a handful of rounds
        return int(min(1 << 29, max(1 << 23, 256 * self.N)))

    def workspace(self, C_, T, cap):
        """One growing scratch buffer shared by all stages (layouts share a common prefix)."""
        nb = C.c_size_t()
        check(lib.nnd_descent_workspace_bytes(self.N, self.k, C_, T, cap, C.byref(nb)))
        if self._ws is None or self._ws.numel() < nb.value:
            self._ws = None
            self._ws = torch.empty(nb.value, dtype=torch.uint8, device=self.device)
        return self._ws

    def all_finite(self):
        """NaN/inf scan of the resident matrix (input validation of check_array, done on the device)."""
        return bool(torch.isfinite(self.X).all())

    def to_host(self, idx_t, dist_t):
        """Device -> host copy of the graph through pinned staging (torch caches the pinned blocks)."""
        hi = torch.empty(idx_t.shape, dtype=idx_t.dtype, pin_memory=True)
        hd = torch.empty(dist_t.shape, dtype=dist_t.dtype, pin_memory=True)
        hi.copy_(idx_t, non_blocking=True); hd.copy_(dist_t, non_blocking=True)
        self.stream.synchronize()
        return hi.numpy(), hd.numpy()

    def launch_count(self):
        return int(lib.nnd_launch_count(self._ctx))

    # ------------------------------------------------------------------ stages
    def make_heap(self):
        check(lib.nnd_set_heap_trusted(self._ctx, 1))
        check(lib.nnd_make_heap(self._ctx, _ptr(self.heap_idx), _ptr(self.heap_dist), _ptr(self.heap_flag), self.N,
                                self.k, self._st()))

    def set_heap(self, idx, dist, flag):
        self.heap_idx.copy_(torch.from_numpy(np.ascontiguousarray(idx, np.int32)))
        self.heap_dist.copy_(torch.from_numpy(np.ascontiguousarray(dist, np.float32)))
        self.heap_flag.copy_(torch.from_numpy(np.ascontiguousarray(flag, np.uint8)))

    def get_heap(self):
        torch.cuda.synchronize(self.device)
        return self.heap_idx.cpu().numpy(), self.heap_dist.cpu().numpy(), self.heap_flag.cpu().numpy()

    def forest(self, tree_rng_states, leaf_size, angular, max_depth=200):
        """-> (leaf_points int32[n_trees*N] device, leaf_offsets int64[n_leaves+1] device, n_leaves, max_leaf)"""
        states = np.ascontiguousarray(tree_rng_states, np.int64).reshape(-1, 3)
        n_trees = states.shape[0]
        nb = C.c_size_t()
        check(lib.nnd_forest_workspace_bytes(self.N, self.d, n_trees, C.byref(nb)))
        ws = torch.empty(nb.value, dtype=torch.uint8, device=self.device)
        leaf_points = torch.empty(n_trees * self.N, dtype=torch.int32, device=self.device)
        leaf_offsets = torch.empty(n_trees * self.N + 1, dtype=torch.int64, device=self.device)
        n_leaves = C.c_int64(); max_leaf = C.c_int32()
        check(lib.nnd_forest_build(self._ctx, _ptr(self.X), self.N, self.d, self.ld, int(bool(angular)),
                                   states.ctypes.data_as(C.POINTER(C.c_int64)), n_trees, int(leaf_size), int(max_depth),
                                   _ptr(leaf_points), _ptr(leaf_offsets), C.byref(n_leaves), C.byref(max_leaf), _ptr(ws),
                                   nb.value, self._st()))
        del ws
        return leaf_points, leaf_offsets[: n_leaves.value + 1], int(n_leaves.value), int(max_leaf.value)

    def leaf_init(self, leaf_points, leaf_offsets, n_leaves, max_leaf, push_capacity=None):
        cap = push_capacity or self.default_push_capacity()
        ws = self.workspace(1, 1, cap)
        npush = C.c_int64()
        check(lib.nnd_leaf_init(self._ctx, C.byref(self.problem()), _ptr(leaf_points), _ptr(leaf_offsets), None,
                                int(n_leaves), int(max_leaf), cap, _ptr(ws), ws.numel(), C.byref(npush), self._st()))
        return int(npush.value)

    def leaf_init_from_array(self, leaf_array, push_capacity=None):
        """Reference-format leaves (rp_trees.py:2918 padded int32[n_leaves,max_leaf], -1 padded)."""
        la = np.ascontiguousarray(leaf_array, np.int32)
        if la.size == 0 or (la.shape == (1, 1) and la[0, 0] < 0):
            return 0
        sizes = (la >= 0).sum(1).astype(np.int64)
        # leaves are left-packed (generate_leaf_updates stops at the first -1)
        first_neg = np.where((la < 0).any(1), (la < 0).argmax(1), la.shape[1])
        sizes = first_neg.astype(np.int64)
        offs = np.zeros(la.shape[0] + 1, np.int64)
        np.cumsum(sizes, out=offs[1:])
        pts = np.concatenate([la[i, : sizes[i]] for i in range(la.shape[0])]) if la.shape[0] < 4096 else \
            la[np.arange(la.shape[1])[None, :] < sizes[:, None]]
        lp = torch.from_numpy(np.ascontiguousarray(pts, np.int32)).to(self.device)
        lo = torch.from_numpy(offs).to(self.device)
        return self.leaf_init(lp, lo, la.shape[0], max(int(sizes.max()), 1), push_capacity)

    def init_random(self, rng_state):
        """rng_state: np.int64[3], advanced in place like the reference does."""
        ws = self._ws if self._ws is not None else self.workspace(1, 1, self.default_push_capacity())
        st = (C.c_int64 * 3)(*[int(v) for v in rng_state])
        check(lib.nnd_init_random(self._ctx, C.byref(self.problem()), st, _ptr(ws), ws.numel(), self._st()))
        rng_state[:] = np.array(list(st), np.int64)

    def init_from_graph(self, init_graph, init_dist=None, push_capacity=None):
        """Warm start (pynndescent_.py:1225-1242, utils.py:836-860): row i receives, in column order, a flagged checked
        push of every init_graph[i, c] >= 0 with d = metric(data[i], data[j]) (nnd_pair_distances) or init_dist[i, c]
        verbatim.  The ordered replay is nnd_apply_records with the column position as emission sequence."""
        dev = self.device
        g = torch.from_numpy(np.ascontiguousarray(init_graph, np.int32)).to(dev)
        N, w = g.shape
        if int(g.max()) >= self.N:
            raise ValueError("init_graph refers to a point outside the dataset")
        cols = g.reshape(-1)
        valid = cols >= 0
        a = torch.arange(N, device=dev, dtype=torch.int32).repeat_interleave(w)[valid].contiguous()
        b = cols[valid].contiguous()
        n = int(a.numel())
        if n == 0:
            return 0
        # caller-supplied distances are pushed verbatim like the reference does: the heap then holds values that are not the metric's
        # own, which is what the presence / duplicate filters of the join and the replay rely on (include/nnd_b200.h)
        check(lib.nnd_set_heap_trusted(self._ctx, 0 if init_dist is not None else 1))
        if init_dist is None:
            d = torch.empty(n, dtype=torch.float32, device=dev)
            check(lib.nnd_pair_distances(self._ctx, C.byref(self.problem()), _ptr(a), _ptr(b), n, _ptr(d), self._st()))
        else:
            d = torch.from_numpy(np.ascontiguousarray(init_dist, np.float32)).to(dev).reshape(-1)[valid].contiguous()
        seq = torch.arange(N * w, device=dev, dtype=torch.int64)[valid].contiguous()
        rec = torch.zeros((n, 6), dtype=torch.int32, device=dev)          # {tgt, oth, d, pad, seq lo, seq hi} = 24 B
        rec[:, 0] = a; rec[:, 1] = b; rec[:, 2] = d.view(torch.int32); rec[:, 4:6] = seq.view(torch.int32).view(n, 2)
        cap = push_capacity or self.default_push_capacity()
        accepted = 0
        if n <= cap:
            accepted = self.apply_records(1, cap, rec.view(torch.uint8).reshape(-1), n)
        else:                                   # records are grouped by target row: cut at row boundaries
            rows_per = max(1, cap // w)
            for r0 in range(0, N, rows_per):
                m = (a >= r0) & (a < r0 + rows_per)
                sub = rec[m].contiguous()
                accepted += self.apply_records(1, cap, sub.view(torch.uint8).reshape(-1), sub.shape[0])
        return accepted

    def _cand_ws(self, max_candidates, n_threads):
        """Candidate scratch (rng tables included); returns (tensor, bytes, fresh) -- fresh = the tables must be regenerated."""
        nb = C.c_size_t()
        check(lib.nnd_cand_workspace_bytes(self.N, self.k, max_candidates, n_threads, C.byref(nb)))
        key = ("cand", max_candidates, n_threads)
        fresh = getattr(self, "_cws_key", None) != key
        if fresh:
            self._cws = None
            self._cws = torch.empty(nb.value, dtype=torch.uint8, device=self.device)
            self._cws_key = key
        return self._cws, nb.value, fresh

    def build_candidates(self, max_candidates, rng_state, n_threads, tables_valid=False):
        _, nbv, fresh = self._cand_ws(max_candidates, n_threads)
        nb = C.c_size_t(nbv)
        if fresh:
            tables_valid = False
        newc = torch.empty((self.N, max_candidates), dtype=torch.int32, device=self.device)
        oldc = torch.empty((self.N, max_candidates), dtype=torch.int32, device=self.device)
        st = (C.c_int64 * 3)(*[int(v) for v in rng_state])
        check(lib.nnd_build_candidates(self._ctx, C.byref(self.problem()), int(max_candidates), st, int(n_threads),
                                       int(bool(tables_valid)), _ptr(newc), _ptr(oldc), _ptr(self._cws), nb.value,
                                       self._st()))
        return newc, oldc

    def process_candidates(self, newc, oldc, push_capacity=None):
        cap = push_capacity or self.default_push_capacity()
        C_ = newc.shape[1]
        ws = self.workspace(C_, 1, cap)
        c = C.c_int64(); stats = (C.c_int64 * 4)()
        check(lib.nnd_process_candidates(self._ctx, C.byref(self.problem()), C_, _ptr(newc), _ptr(oldc), cap, _ptr(ws),
                                         ws.numel(), C.byref(c), stats, self._st()))
        return int(c.value), [int(v) for v in stats]

    def descent(self, max_candidates, rng_state, n_threads, n_iters, delta, push_capacity=None):
        cap = push_capacity or self.default_push_capacity()
        ws = self.workspace(max_candidates, n_threads, cap)
        st = (C.c_int64 * 3)(*[int(v) for v in rng_state])
        iters = C.c_int32(); chist = (C.c_int64 * max(n_iters, 1))(); stats = (C.c_int64 * 4)()
        check(lib.nnd_descent(self._ctx, C.byref(self.problem()), int(max_candidates), st, int(n_threads), int(n_iters),
                              float(delta), cap, _ptr(ws), ws.numel(), C.byref(iters), chist, stats, self._st()))
        return int(iters.value), [int(chist[i]) for i in range(iters.value)], [int(v) for v in stats]

    def deheap_sort(self, correction):
        out = torch.empty_like(self.heap_dist)
        check(lib.nnd_deheap_sort(self._ctx, _ptr(self.heap_idx), _ptr(self.heap_dist), self.N, self.k, int(correction),
                                  _ptr(out), self._st()))
        return self.heap_idx, out

    def build(self, rng_state, tree_rng_states, leaf_size, angular, max_candidates, n_threads, n_iters, delta,
              correction, max_depth=200, push_capacity=None, leaf_array=None, timings=None, init_graph=None,
              init_dist=None):
        """The whole seam-1 + seam-2 pipeline on device-resident data (pynndescent_.py:1118-1260):
        forest -> leaf init -> init_random -> nn_descent_internal -> deheap_sort(+correction).
        Returns (idx tensor, dist tensor, iters, c_hist, stats).  rng_state is advanced in place."""
        import time
        cap = push_capacity or self.default_push_capacity()
        self.workspace(max_candidates, n_threads, cap)
        t0 = time.perf_counter()

        def lap(name):
            nonlocal t0
            if timings is not None:
                self.stream.synchronize()
                t1 = time.perf_counter(); timings[name] = timings.get(name, 0.0) + (t1 - t0); t0 = t1
        self.make_heap()
        if init_graph is not None:
            # nn_descent :339-350: a supplied graph replaces BOTH the leaf init and init_random
            self.workspace(max(1, max_candidates), n_threads, cap)
            self.init_from_graph(init_graph, init_dist, cap)
            lap("init_graph")
            iters, c_hist, stats = self.descent(max_candidates, rng_state, n_threads, n_iters, delta, cap)
            lap("descent")
            idx_t, dist_t = self.deheap_sort(correction)
            lap("deheap_sort")
            return idx_t, dist_t, iters, c_hist, stats
        if leaf_array is not None:
            self.leaf_init_from_array(leaf_array, cap)
        elif tree_rng_states is not None and len(tree_rng_states) > 0:
            lp, lo, n_leaves, max_leaf = self.forest(tree_rng_states, leaf_size, angular, max_depth)
            lap("forest")
            self.leaf_init(lp, lo, n_leaves, max_leaf, cap)
            del lp, lo
        lap("leaf_init")
        self.init_random(rng_state)
        lap("init_random")
        iters, c_hist, stats = self.descent(max_candidates, rng_state, n_threads, n_iters, delta, cap)
        lap("descent")
        idx_t, dist_t = self.deheap_sort(correction)
        lap("deheap_sort")
        return idx_t, dist_t, iters, c_hist, stats

    # ------------------------------------------------------------------ sharded-build primitives (sharded.py)
    def set_row_range(self, lo, hi):
        check(lib.nnd_set_row_range(self._ctx, int(lo), int(hi)))

    def set_thresholds(self, thr):
        """thr: float32 device tensor with the heap roots of ALL rows (kept alive here), or None = snapshot the local heap."""
        self._thr_ext = thr
        check(lib.nnd_set_thresholds(self._ctx, _ptr(thr)))

    def cand_local_counts(self, max_candidates, n_threads, world, owner_block):
        """Step 1 of the row-sharded candidate build: this process's count table, int64[5T + 1 + world]."""
        ws, nb, fresh = self._cand_ws(max_candidates, n_threads)
        self._cand_tables_fresh = fresh or getattr(self, "_cand_tables_fresh", False)
        out = np.zeros(5 * n_threads + 1 + world, np.int64)
        check(lib.nnd_cand_local_counts(self._ctx, C.byref(self.problem()), int(max_candidates), int(n_threads), int(world),
                                        int(owner_block), _ptr(ws), nb, out.ctypes.data_as(C.POINTER(C.c_int64)), self._st()))
        return out

    def cand_export(self, max_candidates, rng_state, n_threads, tables_valid, world, owner_block, glob, n_records):
        """Step 2: forward priorities of the own rows + 12-byte reverse-edge records grouped by owner process (rank order)."""
        ws, nb, fresh = self._cand_ws(max_candidates, n_threads)
        if fresh or getattr(self, "_cand_tables_fresh", False):
            tables_valid = False
        self._cand_tables_fresh = False
        out = torch.empty(max(int(n_records), 1) * 12, dtype=torch.uint8, device=self.device)
        st = (C.c_int64 * 3)(*[int(v) for v in rng_state])
        g = np.ascontiguousarray(glob, np.int64)
        check(lib.nnd_cand_export(self._ctx, C.byref(self.problem()), int(max_candidates), st, int(n_threads), int(bool(tables_valid)),
                                  int(world), int(owner_block), g.ctypes.data_as(C.POINTER(C.c_int64)), _ptr(out), _ptr(ws), nb, self._st()))
        return out

    def cand_import_select(self, max_candidates, n_threads, records, n):
        """Step 3: the received records -> (new, old) candidate lists of the own rows; clears the sampled flags."""
        ws, nb, _ = self._cand_ws(max_candidates, n_threads)
        newc = torch.empty((self.N, max_candidates), dtype=torch.int32, device=self.device)
        oldc = torch.empty((self.N, max_candidates), dtype=torch.int32, device=self.device)
        check(lib.nnd_cand_import_select(self._ctx, C.byref(self.problem()), int(max_candidates), int(n_threads), _ptr(records), int(n),
                                         _ptr(newc), _ptr(oldc), _ptr(ws), nb, self._st()))
        return newc, oldc

    def join_prepare(self, newc, oldc, cap):
        C_ = newc.shape[1]
        ws = self.workspace(C_, 1, cap)
        nr = C.c_int32()
        check(lib.nnd_join_prepare(self._ctx, C.byref(self.problem()), C_, _ptr(newc), _ptr(oldc), cap, _ptr(ws), ws.numel(),
                                   C.byref(nr), self._st()))
        return int(nr.value)

    def join_emit_round(self, C_, cap, rnd):
        ws = self.workspace(C_, 1, cap)
        n = C.c_int64()
        check(lib.nnd_join_emit_round(self._ctx, C.byref(self.problem()), C_, cap, _ptr(ws), ws.numel(), int(rnd), C.byref(n),
                                      self._st()))
        return int(n.value)

    def push_export(self, C_, cap, n_pushed, owner_block, world):
        ws = self.workspace(C_, 1, cap)
        out = torch.empty(max(n_pushed, 1) * 24, dtype=torch.uint8, device=self.device)
        counts = (C.c_int64 * world)()
        check(lib.nnd_push_export(self._ctx, C.byref(self.problem()), C_, cap, _ptr(ws), ws.numel(), int(n_pushed),
                                  int(owner_block), int(world), _ptr(out), counts, self._st()))
        return out, [int(c) for c in counts]

    def apply_records(self, C_, cap, records, n):
        ws = self.workspace(C_, 1, cap)
        c = C.c_int64()
        check(lib.nnd_apply_records(self._ctx, C.byref(self.problem()), C_, _ptr(records), int(n), cap, _ptr(ws), ws.numel(),
                                    C.byref(c), self._st()))
        return int(c.value)

    def leaf_fused_range(self, leaf_points, leaf_offsets, l0, l1, max_leaf, pos0=None, npos=None, cap=None):
        """Leaf join + heap update of leaves [l0,l1) of one tree.  With the position range (pos0, npos) of those leaves and the
        routing workspace size `cap`, the library may use the split leaf init (NND_LEAF_SPLIT) instead of the fused kernel."""
        if pos0 is not None and cap is not None and self._ws is not None:
            check(lib.nnd_leaf_range(self._ctx, C.byref(self.problem()), _ptr(leaf_points), _ptr(leaf_offsets), int(l0), int(l1), int(pos0),
                                     int(npos), int(max_leaf), int(cap), _ptr(self._ws), self._ws.numel(), self._st()))
            return
        check(lib.nnd_leaf_fused_range(self._ctx, C.byref(self.problem()), _ptr(leaf_points), _ptr(leaf_offsets), int(l0), int(l1),
                                       int(max_leaf), self._st()))

    def heap_pack_rows(self, rows, n, oi, od, of, unpack):
        check(lib.nnd_heap_pack_rows(self._ctx, C.byref(self.problem()), _ptr(rows), int(n), _ptr(oi), _ptr(od), _ptr(of),
                                     int(bool(unpack)), self._st()))

    def profile(self, reset=True):
        out = (C.c_double * 10)()
        check(lib.nnd_profile_read(self._ctx, out, int(bool(reset))))
        keys = ("ms", "launches", "rows", "pairs", "pushes")
        return {"join": dict(zip(keys, list(out)[:5])), "leaf": dict(zip(keys, list(out)[5:]))}

    STAGES = ("other", "forest_hyper", "forest_side", "forest_part", "leaf_dist", "leaf_push", "init_random", "cand_fill",
              "cand_select", "join_prep", "join", "scatter", "apply", "deheap", "exchange", "spare")

    def stage_profile(self, enable=True):
        """CUDA-event stage timers inside the library (include/nnd_b200.h NND_ST_*)."""
        check(lib.nnd_stage_profile(self._ctx, int(bool(enable))))

    def stage_ms(self, reset=True):
        out = (C.c_double * len(self.STAGES))()
        check(lib.nnd_stage_read(self._ctx, out, int(bool(reset))))
        return {n: float(v) for n, v in zip(self.STAGES, out) if n != "spare"}

    # ------------------------------------------------------------------ small primitives
    def rng_stream(self, state, start, count, as_float):
        out = torch.empty(count, dtype=torch.float32 if as_float else torch.int32, device=self.device)
        st = (C.c_int64 * 3)(*[int(v) for v in state])
        check(lib.nnd_rng_stream(self._ctx, st, int(start), int(count), int(bool(as_float)), _ptr(out), self._st()))
        torch.cuda.synchronize(self.device)
        return out.cpu().numpy()

    def rng_advance(self, state, n):
        st = (C.c_int64 * 3)(*[int(v) for v in state])
        check(lib.nnd_rng_advance_host(self._ctx, st, int(n)))
        return np.array(list(st), np.int64)

    def pair_distances(self, a, b):
        ta = torch.from_numpy(np.ascontiguousarray(a, np.int32)).to(self.device)
        tb = torch.from_numpy(np.ascontiguousarray(b, np.int32)).to(self.device)
        out = torch.empty(ta.numel(), dtype=torch.float32, device=self.device)
        check(lib.nnd_pair_distances(self._ctx, C.byref(self.problem()), _ptr(ta), _ptr(tb), ta.numel(), _ptr(out),
                                     self._st()))
        torch.cuda.synchronize(self.device)
        return out.cpu().numpy()
