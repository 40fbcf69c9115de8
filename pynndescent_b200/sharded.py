"""Row-sharded NN-descent build across the GPUs of one node (one process per GPU, torch.distributed/NCCL).

Ownership = the reference's own rule for its threads (utils.py:259-264, 706-731): process r owns the heap rows, the candidate
rows and the local-join rows of the contiguous range [r*B, (r+1)*B), B = N // world + 1.  X is replicated (a join on row i
gathers arbitrary vertices); heap tensors keep global row indexing on every process, but only the owner's copy of a row is
authoritative -- the heap is never all-gathered inside the iteration loop.  What crosses NVLink:

  forest      trees t = rank, rank+world, ... are built locally, every tree's leaves (CSR, 4 B per point + 1 B per point for
              "which rank's share holds it") are broadcast
  leaf init   per tree the leaves are split by position among the ranks; the heap rows of a rank's share travel to it from
              the rank that processed them in the previous tree (one all-to-all-v of 9k B per row), and home to their owners
              after the last tree
  candidates  every rank scans ITS rows' edges only (nnd_cand_local_counts / nnd_cand_export / nnd_cand_import_select):
              all-gather of a (5T + 1 + world)-word count table, all-to-all-v of 12-byte reverse-edge proposals
  local join  own rows; thresholds = all-gather of the owners' heap roots (4 B per row); half-updates grouped by
              owner(target), exchanged with one all-to-all-v per round, replayed by the owner in emission-sequence order
  end         all-reduce(sum) of the accepted-push count per iteration; all-gather of the finished rows once

Every half-update carries its global emission sequence and every candidate priority is a closed form of global prefix
counts, so the graph is IDENTICAL to the single-GPU build (and to the oracle) for any world size --
tests/test_sharded_gpu.py, tests/test_parallel_gloo.py and bench.py's `parity` field check exactly that.
"""
import numpy as np
import torch
import torch.distributed as dist

from .engine import Engine


def owner_ranges(n_vertices, world_size):
    """Contiguous vertex ranges exactly as the reference partitions writers (utils.py:259-264)."""
    block = n_vertices // world_size + 1
    return [(min(r * block, n_vertices), min((r + 1) * block, n_vertices)) for r in range(world_size)]


class ShardedBuilder:
    def __init__(self, engine: Engine):
        assert dist.is_initialized(), "torch.distributed must be initialised (backend nccl)"
        self.e = engine
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.N, self.k = engine.N, engine.k
        self.B = self.N // self.world + 1                      # utils.py:259
        self.lo = min(self.rank * self.B, self.N)
        self.hi = min(self.lo + self.B, self.N)
        dev = engine.device
        # padded to world*B rows so that the fixed-size all-gathers (thresholds, finished rows) land in place
        rows = self.world * self.B
        self._full = (torch.empty((rows, self.k), dtype=torch.int32, device=dev),
                      torch.empty((rows, self.k), dtype=torch.float32, device=dev),
                      torch.empty((rows, self.k), dtype=torch.uint8, device=dev))
        engine.heap_idx, engine.heap_dist, engine.heap_flag = (t[: self.N] for t in self._full)
        self._thr = torch.empty(rows, dtype=torch.float32, device=dev)
        self.comm_bytes = 0

    # ------------------------------------------------------------------ small helpers
    def _own(self, t):
        return t[self.rank * self.B: (self.rank + 1) * self.B]

    def _allgather_heap(self, with_flags=True):
        """Owners' row ranges -> everyone (once after the leaf init if init_random has work, and for the finished graph)."""
        for t in self._full[: 3 if with_flags else 2]:
            send = self._own(t).clone()
            dist.all_gather_into_tensor(t, send)
            self.comm_bytes += t.numel() * t.element_size()

    def _allgather_thresholds(self):
        """thr[v] = heap_dist[v, 0] of every vertex from its owner: 4 B per row instead of the whole heap."""
        send = self._own(self._full[1])[:, 0].contiguous()
        dist.all_gather_into_tensor(self._thr, send)
        self.comm_bytes += self._thr.numel() * 4
        self.e.set_thresholds(self._thr)

    def _exchange_rows(self, send_rows, send_counts, recv_rows, recv_counts):
        """all-to-all-v of whole heap rows (idx, dist, flag): `send_rows` grouped by destination rank, `recv_rows` by source rank,
        both in the order the two sides agree on (ascending point id inside a (source, destination) segment)."""
        e, k, dev = self.e, self.k, self.e.device
        ns, nr = int(sum(send_counts)), int(sum(recv_counts))
        bufs = []
        for dt in (torch.int32, torch.float32, torch.uint8):
            bufs.append((torch.empty((max(ns, 1), k), dtype=dt, device=dev), torch.empty((max(nr, 1), k), dtype=dt, device=dev)))
        if ns:
            e.heap_pack_rows(send_rows, ns, bufs[0][0], bufs[1][0], bufs[2][0], False)
        for snd, rcv in bufs:
            dist.all_to_all_single(rcv[:nr].reshape(-1), snd[:ns].reshape(-1), output_split_sizes=[c * k for c in recv_counts],
                                   input_split_sizes=[c * k for c in send_counts])
            self.comm_bytes += nr * k * snd.element_size()
        if nr:
            e.heap_pack_rows(recv_rows, nr, bufs[0][1], bufs[1][1], bufs[2][1], True)

    def _move_rows(self, have, dest_of_have, want, src_of_want):
        """Bring the heap rows of the points `want` (int32 device tensor: what this rank works on next) up to date.

        `have` = the points whose current row copy is here, `dest_of_have` = the rank that works on each of them next,
        `src_of_want` = the rank that holds the current copy of each wanted point (all derived from replicated plan data, so
        the two sides of every (source, destination) pair list the same points: ascending point id).  Rows that stay are not
        moved; one host sync (the two count vectors)."""
        W, me, N = self.world, self.rank, self.N
        go = dest_of_have != me
        s_pts, s_dst = have[go], dest_of_have[go]
        s_pts = s_pts[torch.argsort(s_dst.to(torch.int64) * N + s_pts.to(torch.int64))].contiguous()
        come = src_of_want != me
        r_pts, r_src = want[come], src_of_want[come]
        r_pts = r_pts[torch.argsort(r_src.to(torch.int64) * N + r_pts.to(torch.int64))].contiguous()
        counts = torch.stack([torch.bincount(s_dst.long(), minlength=W), torch.bincount(r_src.long(), minlength=W)]).cpu().numpy()
        self._exchange_rows(s_pts, [int(c) for c in counts[0]], r_pts, [int(c) for c in counts[1]])

    # ------------------------------------------------------------------ stages
    def forest_and_leaf_init(self, tree_states, leaf_size, angular, max_depth, cap):
        e, N, dev = self.e, self.N, self.e.device
        W, me = self.world, self.rank
        n_trees = len(tree_states)
        mine = [t for t in range(n_trees) if t % W == me]
        meta = torch.zeros((n_trees, 2), dtype=torch.int64, device=dev)          # (n_leaves, max_leaf) per tree
        local = {}
        if mine:
            lp, lo, n_leaves, max_leaf = e.forest(tree_states[mine], leaf_size, angular, max_depth)
            lo_h = lo.cpu().numpy()
            for j, t in enumerate(mine):                     # split the tree-major CSR back into trees
                a = int(np.searchsorted(lo_h, j * N)); b = int(np.searchsorted(lo_h, (j + 1) * N))
                local[t] = (lp[j * N: (j + 1) * N], lo[a: b + 1] - j * N, lo_h[a: b + 1] - j * N)
                meta[t, 0] = b - a; meta[t, 1] = max_leaf
        dist.all_reduce(meta, op=dist.ReduceOp.SUM)
        meta_h = meta.cpu().numpy()
        max_leaf = int(meta_h[:, 1].max())

        def cuts_of(oh, n_leaves_t):
            """rank r takes the leaves whose first position lies in [r*N/world, (r+1)*N/world): (leaf cuts, position cuts)"""
            cuts = [int(np.searchsorted(oh[:-1], (r * N) // W)) for r in range(W)] + [int(n_leaves_t)]
            return cuts, [int(oh[c]) for c in cuts]
        # every tree's leaves to every rank: one points buffer + one offsets buffer, filled by broadcasts; with them travels
        # share_of[t][p] = the rank whose share of tree t contains point p (one N-sized scatter on the tree's owner only)
        pts_all = torch.empty((n_trees, N), dtype=torch.int32, device=dev)
        share_of = torch.empty((n_trees, N), dtype=torch.uint8, device=dev)
        offs_all = torch.empty(int(meta_h[:, 0].sum()) + n_trees, dtype=torch.int64, device=dev)
        obase = np.concatenate([[0], np.cumsum(meta_h[:, 0] + 1)])
        ar = None
        for t in range(n_trees):
            owner = t % W
            ov = offs_all[obase[t]: obase[t + 1]]
            if owner == me:
                pts_all[t].copy_(local[t][0]); ov.copy_(local[t][1])
                _, p0 = cuts_of(local[t][2], meta_h[t, 0])
                if ar is None:
                    ar = torch.arange(N, dtype=torch.int64, device=dev)
                bounds = torch.tensor(p0[1:-1], dtype=torch.int64, device=dev)
                share_of[t][local[t][0].long()] = torch.bucketize(ar, bounds, right=True).to(torch.uint8)
            dist.broadcast(pts_all[t], src=owner)
            dist.broadcast(ov, src=owner)
            dist.broadcast(share_of[t], src=owner)
        del ar
        self.comm_bytes += pts_all.numel() * 5 + offs_all.numel() * 8
        offs_h = offs_all.cpu().numpy()
        plan = [cuts_of(offs_h[obase[t]: obase[t + 1]], meta_h[t, 0]) for t in range(n_trees)]
        # Before the first tree every rank holds the same fresh heap (make_heap), so nothing moves for tree 0; afterwards the
        # current copy of row p sits on rank share_of[t][p] of the tree processed last.
        fused_ok = True
        prev = None                                                               # (tree index, my share of it)
        for t in range(n_trees):
            cuts, p0 = plan[t]
            pts = pts_all[t]; offs = offs_all[obase[t]: obase[t + 1]]
            share = pts[p0[me]: p0[me + 1]]
            if prev is not None:
                self._move_rows(prev[1], share_of[t][prev[1].long()].to(torch.int32), share, share_of[prev[0]][share.long()].to(torch.int32))
            try:
                e.leaf_fused_range(pts, offs, cuts[me], cuts[me + 1], max_leaf, pos0=p0[me], npos=p0[me + 1] - p0[me], cap=cap)
            except NotImplementedError:
                fused_ok = False
                break
            prev = (t, share)
        if not fused_ok:
            # leaves too large for the fused / split kernels: every rank runs the routed leaf init on all trees (replicated)
            e.make_heap()
            lo = torch.cat([offs_all[obase[t]: obase[t + 1] - 1] + t * N for t in range(n_trees)] +
                           [torch.tensor([n_trees * N], dtype=torch.int64, device=dev)])
            e.leaf_init(pts_all.reshape(-1), lo, lo.numel() - 1, max_leaf, cap)
            return
        if prev is not None:
            # rows go home: the owner of p is p // B
            own = torch.arange(self.lo, self.hi, dtype=torch.int32, device=dev)
            self._move_rows(prev[1], torch.clamp(prev[1].to(torch.int64) // self.B, max=W - 1).to(torch.int32), own,
                            share_of[prev[0]][self.lo: self.hi].to(torch.int32))

    def init_random(self, rng_state):
        """pynndescent_.py:188-203 on the sharded heap: rows whose root is empty draw from ONE sequential stream in row order.
        After an RP-forest init that is no row at all; only when some row needs draws are the rows made consistent everywhere
        (one all-gather) and the replicated kernel run -- identical results, the common case costs one 8-byte all-reduce."""
        e = self.e
        own = e.heap_idx[self.lo: self.hi]
        need = torch.where(own[:, 0] < 0, (own < 0).sum(1), torch.zeros((), dtype=torch.int64, device=e.device)).sum().reshape(1)
        dist.all_reduce(need, op=dist.ReduceOp.SUM)
        if int(need[0]) == 0:
            return
        self._allgather_heap()
        e.init_random(rng_state)

    def candidates(self, C_, rng_state, T, tables_valid):
        """new_build_candidates for the own rows: count table all-gather + all-to-all-v of the reverse-edge proposals."""
        e, W, me = self.e, self.world, self.rank
        loc = e.cand_local_counts(C_, T, W, self.B)                               # int64[5T + 1 + W]
        L = loc.shape[0]
        allt = torch.empty(W * L, dtype=torch.int64, device=e.device)
        dist.all_gather_into_tensor(allt, torch.from_numpy(loc).to(e.device))
        A = allt.cpu().numpy().reshape(W, L)
        glob = np.zeros(6 * T, np.int64)
        for n in range(T):
            glob[6 * n: 6 * n + 4] = A[:, 5 * n: 5 * n + 4].sum(0)               # R_s, V_s, V_e, R_e over all processes
            glob[6 * n + 4] = A[:me, 5 * n + 4].sum()                            # lower ranks' edges into R_n
            glob[6 * n + 5] = A[:me, 5 * T].sum()                                # lower ranks' valid edges
        send_counts = [int(c) for c in A[me, 5 * T + 1:]]
        recv_counts = [int(A[q, 5 * T + 1 + me]) for q in range(W)]
        ns, nr = sum(send_counts), sum(recv_counts)
        send = e.cand_export(C_, rng_state, T, tables_valid, W, self.B, glob, ns)
        recv = torch.empty(max(nr, 1) * 12, dtype=torch.uint8, device=e.device)
        dist.all_to_all_single(recv[: nr * 12], send[: ns * 12], output_split_sizes=[c * 12 for c in recv_counts],
                               input_split_sizes=[c * 12 for c in send_counts])
        self.comm_bytes += nr * 12 + L * 8 * W
        return e.cand_import_select(C_, T, recv, nr)

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
        e.set_row_range(self.lo, self.hi)
        try:
            for it in range(n_iters):
                newc, oldc = self.candidates(C_, rng_state, T, tables_valid=(it > 0))
                lap("d.candidates")
                self._allgather_thresholds()
                my_rounds = e.join_prepare(newc, oldc, cap)
                lap("d.join_prepare")
                # Canonical push order per target row = ascending emission sequence.  All ranks emit their round r at
                # the same time; the owner only COLLECTS the records of every round and replays them once at the end
                # of the iteration, where the sort by emission sequence restores the canonical order no matter which
                # rank / round a record came from.  (Emission only reads the threshold snapshot, which does not change
                # inside an iteration, so nothing depends on the replay having happened.)  Ranks may need different
                # numbers of rounds (and a round that overflows the push list is split on the spot): the count exchange
                # carries a "more rounds follow" flag and the loop ends when no rank raises it.
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
                ct = torch.tensor([c_loc], dtype=torch.int64, device=e.device)
                dist.all_reduce(ct, op=dist.ReduceOp.SUM)
                c_hist.append(int(ct[0]))
                if c_hist[-1] <= delta * self.k * self.N:
                    break
        finally:
            e.set_row_range(0, 2 ** 62)
            e.set_thresholds(None)
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
        self.init_random(rng_state)
        iters, c_hist = self.descent(C_, rng_state, T, n_iters, delta, cap, timings=timings)
        lap("descent")
        # the finished rows to everyone, then the heap-sort tail + correction on all rows (0.2 ms per million rows)
        self._allgather_heap(with_flags=False)
        idx_t, dist_t = e.deheap_sort(correction)
        lap("deheap_sort")
        return idx_t, dist_t, iters, c_hist
