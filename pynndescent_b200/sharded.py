"""Row-sharded NN-descent build across the GPUs of one node (one process per GPU, torch.distributed/NCCL).

Every rank keeps X and the WHOLE heap (N x k x 9 B -- 1.35 GB at 10M x 15), but works only on its own
contiguous row range [r*B, (r+1)*B), B = N // world + 1 -- the reference's own ownership rule for its
threads (utils.py:259-264):

  forest      trees t = rank, rank+world, ... are built locally, then every tree's leaves are broadcast
  leaf init   per tree, the leaves are split by position among the ranks (disjoint point sets), each rank runs
              the fused leaf join on its share, then the updated heap rows are all-gathered
  candidates  sampled for the own rows only (reverse edges are found by scanning the replicated heap)
  local join  own rows; the half-updates are grouped by owner(target) and exchanged with ONE all-to-all-v per
              round; the owner replays them in emission-sequence order
  end of it.  all-gather of the owners' heap rows, all-reduce(sum) of the accepted-push count `c`

Every half-update carries its global emission sequence, so the graph is IDENTICAL to the single-GPU build
(and to the oracle) for any world size -- tests/test_sharded_gpu.py checks exactly that.
"""
import os

import numpy as np
import torch
import torch.distributed as dist

from .engine import Engine


class ShardedBuilder:
    def __init__(self, engine: Engine):
        assert dist.is_initialized(), "torch.distributed must be initialised (backend nccl)"
        self.e = engine
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.N, self.k = engine.N, engine.k
        self.B = self.N // self.world + 1                      # utils.py:259
        self.lo = min(self.rank * self.B, self.N)
        self.hi = min(self.lo + self.B, self.N)
        # heap tensors with room for world*B rows so that the per-range all-gather lands in place
        dev = engine.device
        rows = self.world * self.B
        self._full = (torch.empty((rows, self.k), dtype=torch.int32, device=dev),
                      torch.empty((rows, self.k), dtype=torch.float32, device=dev),
                      torch.empty((rows, self.k), dtype=torch.uint8, device=dev))
        engine.heap_idx, engine.heap_dist, engine.heap_flag = (t[: self.N] for t in self._full)
        self.comm_bytes = 0

    # ------------------------------------------------------------------ collectives
    def _allgather_heap(self):
        """Owners' row ranges -> everyone (3 all-gathers of B x k rows)."""
        for t in self._full:
            send = t[self.rank * self.B: (self.rank + 1) * self.B].clone()
            dist.all_gather_into_tensor(t, send)
            self.comm_bytes += t.numel() * t.element_size()

    # ------------------------------------------------------------------ stages
    def forest_and_leaf_init(self, tree_states, leaf_size, angular, max_depth, cap):
        e, N, k, dev = self.e, self.N, self.k, self.e.device
        n_trees = len(tree_states)
        mine = [t for t in range(n_trees) if t % self.world == self.rank]
        meta = torch.zeros((n_trees, 2), dtype=torch.int64, device=dev)          # (n_leaves, max_leaf) per tree
        local = {}
        if mine:
            lp, lo, n_leaves, max_leaf = e.forest(tree_states[mine], leaf_size, angular, max_depth)
            lo_h = lo.cpu().numpy()
            for j, t in enumerate(mine):                     # split the tree-major CSR back into trees
                a = int(np.searchsorted(lo_h, j * N)); b = int(np.searchsorted(lo_h, (j + 1) * N))
                local[t] = (lp[j * N: (j + 1) * N], lo[a: b + 1] - j * N)
                meta[t, 0] = b - a; meta[t, 1] = max_leaf
        dist.all_reduce(meta, op=dist.ReduceOp.SUM)
        meta_h = meta.cpu().numpy()
        max_leaf = int(meta_h[:, 1].max())
        # every tree's leaves to every rank: one points buffer + one offsets buffer, filled by broadcasts
        pts_all = torch.empty((n_trees, N), dtype=torch.int32, device=dev)
        offs_all = torch.empty(int(meta_h[:, 0].sum()) + n_trees, dtype=torch.int64, device=dev)
        obase = np.concatenate([[0], np.cumsum(meta_h[:, 0] + 1)])
        for t in range(n_trees):
            owner = t % self.world
            ov = offs_all[obase[t]: obase[t + 1]]
            if owner == self.rank:
                pts_all[t].copy_(local[t][0]); ov.copy_(local[t][1])
            dist.broadcast(pts_all[t], src=owner)
            dist.broadcast(ov, src=owner)
        self.comm_bytes += pts_all.numel() * 4 + offs_all.numel() * 8
        offs_h = offs_all.cpu().numpy()
        # rank r takes, in every tree, the leaves whose first position lies in [r*N/world, (r+1)*N/world)
        plan = []
        for t in range(n_trees):
            oh = offs_h[obase[t]: obase[t + 1]]
            cuts = [int(np.searchsorted(oh[:-1], (r * N) // self.world)) for r in range(self.world)] + [int(meta_h[t, 0])]
            plan.append((cuts, [int(oh[c]) for c in cuts]))
        maxn = max(p0[r + 1] - p0[r] for _, p0 in plan for r in range(self.world))
        si = torch.zeros((maxn, k), dtype=torch.int32, device=dev); sd = torch.zeros((maxn, k), dtype=torch.float32, device=dev)
        sf = torch.zeros((maxn, k), dtype=torch.uint8, device=dev)
        gi = torch.empty((self.world * maxn, k), dtype=torch.int32, device=dev)
        gd = torch.empty((self.world * maxn, k), dtype=torch.float32, device=dev)
        gf = torch.empty((self.world * maxn, k), dtype=torch.uint8, device=dev)
        # NND_LEAF_EXCHANGE=targeted (default off until run over NCCL; the CPU gloo tests cover both modes): between two trees a
        # rank only needs the rows of the points it is about to process, and it knows (everything here is replicated plan data)
        # which rank processed each of them in the previous tree -- so every rank sends every other rank exactly those rows with
        # one all-to-all-v per tree (N/world rows in, N/world out) instead of all-gathering all N rows; the replicas are made
        # consistent by the ordinary all-gather after the LAST tree, whose shares contain every point exactly once.
        targeted = os.environ.get("NND_LEAF_EXCHANGE", "") == "targeted" and self.world > 1
        fused_ok = True
        for t in range(n_trees):
            cuts, p0 = plan[t]
            pts = pts_all[t]; offs = offs_all[obase[t]: obase[t + 1]]
            if targeted and t > 0:
                self._fetch_rows(pts_all[t - 1], plan[t - 1][1], pts, p0, maxn)
            try:
                e.leaf_fused_range(pts, offs, cuts[self.rank], cuts[self.rank + 1], max_leaf, pos0=p0[self.rank],
                                   npos=p0[self.rank + 1] - p0[self.rank], cap=cap)
            except NotImplementedError:
                fused_ok = False
                break
            if targeted and t + 1 < n_trees:
                continue
            # exchange the rows each rank has just updated (its leaves' points, in position order)
            e.heap_pack_rows(pts[p0[self.rank]: p0[self.rank + 1]], p0[self.rank + 1] - p0[self.rank], si, sd, sf, False)
            dist.all_gather_into_tensor(gi, si); dist.all_gather_into_tensor(gd, sd); dist.all_gather_into_tensor(gf, sf)
            self.comm_bytes += gi.numel() * 9
            for r in range(self.world):
                if r != self.rank:
                    e.heap_pack_rows(pts[p0[r]: p0[r + 1]], p0[r + 1] - p0[r], gi[r * maxn:], gd[r * maxn:], gf[r * maxn:], True)
        if not fused_ok:
            # leaves too large for the fused kernel: every rank runs the routed leaf init on all trees (replicated)
            e.make_heap()
            lo = torch.cat([offs_all[obase[t]: obase[t + 1] - 1] + t * N for t in range(n_trees)] +
                           [torch.tensor([n_trees * N], dtype=torch.int64, device=dev)])
            e.leaf_init(pts_all.reshape(-1), lo, lo.numel() - 1, max_leaf, cap)

    def _fetch_rows(self, prev_pts, prev_p0, pts, p0, maxn):
        """Bring the heap rows of the points this rank processes in the coming tree (positions [p0[rank], p0[rank+1]) of `pts`)
        up to date: each was last written by the rank whose share of the PREVIOUS tree (`prev_pts`, `prev_p0`) contains it."""
        e, k, dev, W, me = self.e, self.k, self.e.device, self.world, self.rank
        N = self.N
        pos_prev = torch.empty(N, dtype=torch.int64, device=dev)
        pos_prev[prev_pts.long()] = torch.arange(N, dtype=torch.int64, device=dev)
        bounds = torch.tensor(prev_p0[1:-1], dtype=torch.int64, device=dev)
        writer = torch.bucketize(pos_prev, bounds, right=True)                       # rank that processed each point last
        send_rows, send_counts, recv_rows, recv_counts = [], [], [], []
        for q in range(W):
            theirs = pts[p0[q]: p0[q + 1]]                                            # what rank q is about to process ...
            mine_for_q = theirs[writer[theirs.long()] == me] if q != me else theirs[:0]   # ... and I wrote last
            send_rows.append(mine_for_q); send_counts.append(int(mine_for_q.numel()))
            my_next = pts[p0[me]: p0[me + 1]]
            from_q = my_next[writer[my_next.long()] == q] if q != me else my_next[:0]
            recv_rows.append(from_q); recv_counts.append(int(from_q.numel()))
        ns, nr = sum(send_counts), sum(recv_counts)
        srows = torch.cat(send_rows) if ns else torch.zeros(0, dtype=torch.int32, device=dev)
        bufs = []
        for dt in (torch.int32, torch.float32, torch.uint8):
            bufs.append((torch.zeros((max(ns, 1), k), dtype=dt, device=dev), torch.zeros((max(nr, 1), k), dtype=dt, device=dev)))
        if ns:
            e.heap_pack_rows(srows.contiguous(), ns, bufs[0][0], bufs[1][0], bufs[2][0], False)
        for snd, rcv in bufs:
            dist.all_to_all_single(rcv[:nr].reshape(-1), snd[:ns].reshape(-1), output_split_sizes=[c * k for c in recv_counts],
                                   input_split_sizes=[c * k for c in send_counts])
            self.comm_bytes += nr * k * snd.element_size()
        if nr:
            e.heap_pack_rows(torch.cat(recv_rows).contiguous(), nr, bufs[0][1], bufs[1][1], bufs[2][1], True)

    def _replay(self, C_, cap, inbox):
        """Apply the collected 24-byte records to the own heap rows; returns the accepted-push count.

        One call when they fit the routing workspace (`cap` records).  Otherwise the own row range is halved
        recursively and each half replayed on its own -- pushes to different rows commute, so any split BY
        TARGET ROW keeps the result identical.
        """
        e = self.e
        if not inbox:
            return 0
        buf = inbox[0] if len(inbox) == 1 else torch.cat(inbox)
        n = buf.numel() // 24
        if n <= cap:
            return e.apply_records(C_, cap, buf, n)
        recs = buf.view(n, 24)
        tgt = buf.view(torch.int32).view(n, 6)[:, 0]
        c = 0
        work = [(self.lo, self.hi, recs, tgt)]
        while work:
            a, b, r, t = work.pop()
            m = r.shape[0]
            if m == 0:
                continue
            if m <= cap:
                c += e.apply_records(C_, cap, r.reshape(-1), m)
                continue
            if b - a <= 1:
                raise OverflowError("push_capacity %d is smaller than the %d half-updates of ONE heap row" % (cap, m))
            mid = (a + b) // 2
            left = t < mid
            rl = r[left].contiguous(); rr = r[~left].contiguous()
            work.append((mid, b, rr, rr.view(-1).view(torch.int32).view(-1, 6)[:, 0]))
            work.append((a, mid, rl, rl.view(-1).view(torch.int32).view(-1, 6)[:, 0]))
        return c

    def descent(self, C_, rng_state, T, n_iters, delta, cap, timings=None):
        import time
        e = self.e
        c_hist = []
        tl = [time.perf_counter()]

        def lap(name):
            if timings is not None:
                e.stream.synchronize()
                t1 = time.perf_counter(); timings[name] = timings.get(name, 0.0) + (t1 - tl[0]); tl[0] = t1
        for it in range(n_iters):
            e.set_row_range(self.lo, self.hi)
            newc, oldc = e.build_candidates(C_, rng_state, T, tables_valid=(it > 0))
            lap("d.candidates")
            my_rounds = e.join_prepare(newc, oldc, cap)
            lap("d.join_prepare")
            # Canonical push order per target row = ascending emission sequence.  All ranks emit their round r at
            # the same time; the owner only COLLECTS the records of every round and replays them once at the end
            # of the iteration, where the sort by emission sequence restores the canonical order no matter which
            # rank / round a record came from.  (Emission only reads the replicated heap's thresholds, which do
            # not change before the all-gather, so nothing depends on the replay having happened.)  Ranks may
            # need different numbers of rounds (and a round that overflows the push list is split on the spot):
            # the count exchange carries a "more rounds follow" flag and the loop ends when no rank raises it.
            inbox = []
            rnd = 0
            while True:
                n = 0
                while rnd < my_rounds:
                    n = e.join_emit_round(C_, cap, rnd)
                    if n >= 0:
                        break
                    my_rounds += 1                                   # round split in two: emit its first half
                lap("d.join_emit")
                send, counts = e.push_export(C_, cap, n, self.B, self.world)
                lap("d.export")
                more = 1 if rnd + 1 < my_rounds else 0
                cs = torch.tensor([[c, more] for c in counts], dtype=torch.int64, device=e.device)
                cr = torch.empty_like(cs)
                dist.all_to_all_single(cr, cs)
                crh = cr.cpu().numpy()
                rc = [int(v) for v in crh[:, 0]]
                recv = torch.empty(max(sum(rc), 1) * 24, dtype=torch.uint8, device=e.device)
                dist.all_to_all_single(recv[: sum(rc) * 24], send[: n * 24], output_split_sizes=[c * 24 for c in rc],
                                       input_split_sizes=[c * 24 for c in counts])
                self.comm_bytes += sum(rc) * 24
                if sum(rc) > 0:
                    inbox.append(recv[: sum(rc) * 24])
                lap("d.all_to_all")
                rnd += 1
                if not crh[:, 1].any():
                    break
            c_loc = self._replay(C_, cap, inbox)
            lap("d.apply")
            e.set_row_range(0, 2 ** 62)
            self._allgather_heap()
            lap("d.allgather_heap")
            ct = torch.tensor([c_loc], dtype=torch.int64, device=e.device)
            dist.all_reduce(ct, op=dist.ReduceOp.SUM)
            c_hist.append(int(ct[0]))
            if c_hist[-1] <= delta * self.k * self.N:
                break
        return len(c_hist), c_hist

    def build(self, rng_state, tree_states, leaf_size, angular, C_, T, n_iters, delta, correction, max_depth=200, cap=None,
              timings=None):
        import time
        e = self.e
        cap = cap or e.default_push_capacity()
        e.workspace(C_, T, cap)
        t0 = time.perf_counter()

        def lap(name):
            nonlocal t0
            if timings is not None:
                e.stream.synchronize()
                t1 = time.perf_counter(); timings[name] = timings.get(name, 0.0) + (t1 - t0); t0 = t1
        e.make_heap()
        if tree_states is not None and len(tree_states) > 0:
            self.forest_and_leaf_init(np.asarray(tree_states), leaf_size, angular, max_depth, cap)
        lap("forest+leaf_init")
        e.init_random(rng_state)
        iters, c_hist = self.descent(C_, rng_state, T, n_iters, delta, cap, timings=timings)
        lap("descent")
        idx_t, dist_t = e.deheap_sort(correction)
        lap("deheap_sort")
        return idx_t, dist_t, iters, c_hist
