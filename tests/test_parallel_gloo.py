"""CPU-only: the N>1 host logic over the gloo backend, world_size 2 (rendezvous on 127.0.0.1)."""
import os
import socket
import subprocess
import sys

import numpy as np

from conftest import ROOT

def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def test_owner_ranges_match_reference_rule():
    from pynndescent_b200.sharded import owner_ranges
    for n, w in ((1000, 8), (1001, 2), (17, 4), (10_000_000, 8)):
        rr = owner_ranges(n, w)
        assert rr[0][0] == 0 and rr[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(rr, rr[1:]))
        block = n // w + 1                                  # utils.py:259
        assert all(hi - lo <= block for lo, hi in rr)


SHARDED_WORKER = r'''
"""Drive pynndescent_b200.sharded.ShardedBuilder.descent over gloo with a scripted engine: checks the exchange
protocol only (unequal round counts, rounds split on overflow, the "more rounds" flag, the replay split by
target row) -- every emitted record must reach the owner of its target exactly once."""
import os, sys
sys.path.insert(0, os.environ["REPO_ROOT"])
import numpy as np, torch, torch.distributed as dist
dist.init_process_group("gloo")
from pynndescent_b200.sharded import ShardedBuilder
r, w = dist.get_rank(), dist.get_world_size()
N, K, CAP = 1001, 4, 64


class ScriptedEngine:
    """Stands in for engine.Engine: emits seeded records instead of running the CUDA join."""
    def __init__(self):
        self.N, self.k, self.device = N, K, torch.device("cpu")
        self.heap_idx = self.heap_dist = self.heap_flag = None
        self.emitted, self.received, self.it = [], [], 0
        self.lo = self.hi = None
    def set_row_range(self, lo, hi): self.lo, self.hi = lo, hi
    def set_thresholds(self, thr): self.thr = thr
    def cand_local_counts(self, C_, T, world, B): return np.zeros(5 * T + 1 + world, np.int64)
    def cand_export(self, C_, rng, T, tables_valid, world, B, glob, n): return torch.zeros(12, dtype=torch.uint8)
    def cand_import_select(self, C_, T, records, n): return None, None
    def join_prepare(self, newc, oldc, cap):
        self.it += 1
        rs = np.random.RandomState(100 * self.it + r)
        # rank 0 plans 3 rounds, rank 1 one round (iteration 2: rank 0 none at all)
        n_rounds = (3 if r == 0 else 1) if self.it == 1 else (0 if r == 0 else 2)
        self.rounds = [rs.randint(0, N, size=rs.randint(20, 60)) for _ in range(n_rounds)]
        self.split_pending = (self.it == 1 and r == 0)
        return n_rounds
    def join_emit_round(self, C_, cap, rnd):
        if rnd >= len(self.rounds): return 0
        if self.split_pending and rnd == 1:           # this round "overflows" once: split in two, caller retries
            self.split_pending = False
            t = self.rounds[rnd]; self.rounds[rnd: rnd + 1] = [t[: len(t) // 2], t[len(t) // 2:]]
            return -1
        self.cur = self.rounds[rnd]
        return len(self.cur)
    def push_export(self, C_, cap, n, B, world):
        if n == 0: return torch.zeros(24, dtype=torch.uint8), [0] * world
        tgt = np.sort(self.cur.astype(np.int32), kind="stable")          # grouped by owner since owners own ranges
        rec = np.zeros((n, 6), np.int32)
        rec[:, 0] = tgt; rec[:, 1] = r; rec[:, 4] = np.arange(n) + 1000 * len(self.emitted); rec[:, 5] = self.it
        self.emitted.append(rec.copy())
        counts = [int(((tgt // B).clip(max=world - 1) == o).sum()) for o in range(world)]
        return torch.from_numpy(rec).view(torch.uint8).reshape(-1), counts
    def apply_records(self, C_, cap, buf, n):
        assert n <= cap, "replay must split by row below the capacity"
        rec = buf[: n * 24].view(torch.int32).view(n, 6).numpy()
        assert ((rec[:, 0] >= self.lo) & (rec[:, 0] < self.hi)).all(), "record delivered to the wrong owner"
        self.received.append(rec.copy())
        return n


e = ScriptedEngine()
sb = ShardedBuilder(e)
for t in sb._full:
    t.zero_()
iters, c_hist = sb.descent(4, None, 1, 2, -1.0, CAP)
sent = np.concatenate(e.emitted) if e.emitted else np.zeros((0, 6), np.int32)
got = np.concatenate(e.received) if e.received else np.zeros((0, 6), np.int32)
allsent, allgot = [None] * w, [None] * w
dist.all_gather_object(allsent, sent); dist.all_gather_object(allgot, got)
if r == 0:
    S = np.concatenate(allsent); G = np.concatenate(allgot)
    key = lambda a: sorted(map(tuple, a.tolist()))
    ok = key(S) == key(G) and iters == 2 and c_hist == [len(np.concatenate([x[x[:, 5] == i] for x in allsent])) for i in (1, 2)]
    print("RESULT", "OK" if ok else "BAD", len(S), len(G), c_hist, flush=True)
dist.destroy_process_group()
'''


def test_sharded_exchange_protocol_gloo(tmp_path):
    """world_size 2 over gloo: ShardedBuilder.descent's round / flag / replay logic with a scripted engine."""
    script = tmp_path / "sharded_worker.py"
    script.write_text(SHARDED_WORKER)
    env = dict(os.environ, REPO_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][0].split()
    assert line[1] == "OK", out.stdout[-2000:]


LEAF_WORKER = r'''
"""ShardedBuilder.forest_and_leaf_init over gloo with an oracle-backed engine: trees built by their owner rank, leaves broadcast,
every tree's leaves split among the ranks, the heap rows of a rank's share fetched from whoever processed them in the previous
tree, rows sent home after the last tree.  Afterwards every rank must hold, for ITS row range, exactly the heap the
oracle's init_rp_tree produces from the whole forest -- and every leaf_fused_range call must carry the position range of its
leaves (what nnd_leaf_range needs for the split leaf init)."""
import os, sys
sys.path.insert(0, os.environ["REPO_ROOT"]); sys.path.insert(0, os.path.join(os.environ["REPO_ROOT"], "tests"))
import numpy as np, torch, torch.distributed as dist
dist.init_process_group("gloo")
from conftest import latent
from oracle import nnd_oracle as orc
from pynndescent_b200.sharded import ShardedBuilder
r, w = dist.get_rank(), dist.get_world_size()
N, K, LEAF, NT = 600, 8, 20, 5
X = latent(N, 12, 2)
states = np.random.RandomState(7).randint(-2 ** 31 + 1, 2 ** 31 - 1, size=(NT, 3)).astype(np.int64)


def csr(la):
    sizes = (la >= 0).sum(1)
    offs = np.zeros(la.shape[0] + 1, np.int64); np.cumsum(sizes, out=offs[1:])
    return la[la >= 0].astype(np.int32), offs, int(sizes.max())


class OracleShardEngine:
    def __init__(self):
        self.N, self.k, self.device = N, K, torch.device("cpu")
        self.heap_idx = self.heap_dist = self.heap_flag = None
        self.calls = []
    def forest(self, tree_states, leaf_size, angular, max_depth):
        la = orc.forest_leaf_array(X, np.ascontiguousarray(tree_states, np.int64), leaf_size, angular, max_depth, "level")
        pts, offs, ml = csr(la)
        return torch.from_numpy(pts), torch.from_numpy(offs), la.shape[0], ml
    def leaf_fused_range(self, pts, offs, l0, l1, max_leaf, pos0=None, npos=None, cap=None):
        o = offs.numpy(); p = pts.numpy()
        assert pos0 == int(o[l0]) and npos == int(o[l1] - o[l0]) and cap is not None
        self.calls.append((l0, l1))
        if l1 == l0: return
        la = np.full((l1 - l0, max_leaf), -1, np.int32)
        for i, l in enumerate(range(l0, l1)):
            m = p[o[l]: o[l + 1]]; la[i, : len(m)] = m
        heap = (self.heap_idx.numpy(), self.heap_dist.numpy(), self.heap_flag.numpy())
        orc.init_rp_tree(X, "euclidean", heap, la)
    def heap_pack_rows(self, rows, n, oi, od, of, unpack):
        if n == 0: return
        idx = rows[:n].long()
        if not unpack:
            oi[:n] = self.heap_idx[idx]; od[:n] = self.heap_dist[idx]; of[:n] = self.heap_flag[idx]
        else:
            self.heap_idx[idx] = oi[:n]; self.heap_dist[idx] = od[:n]; self.heap_flag[idx] = of[:n]


e = OracleShardEngine()
sb = ShardedBuilder(e)
hi, hd, hf = orc.make_heap(N, K)
e.heap_idx.copy_(torch.from_numpy(hi)); e.heap_dist.copy_(torch.from_numpy(hd)); e.heap_flag.copy_(torch.from_numpy(hf))
sb.forest_and_leaf_init(states, LEAF, False, 200, 1 << 20)
ref = orc.make_heap(N, K)
orc.init_rp_tree(X, "euclidean", ref, orc.forest_leaf_array(X, states, LEAF, False, 200, "level"))
lo, hi = sb.lo, sb.hi            # after the sharded leaf init every rank holds the finished rows of ITS range (rows travel home)
ok = (np.array_equal(e.heap_idx.numpy()[lo:hi], ref[0][lo:hi]) and np.array_equal(e.heap_dist.numpy()[lo:hi], ref[1][lo:hi])
      and np.array_equal(e.heap_flag.numpy()[lo:hi], ref[2][lo:hi]) and len(e.calls) == NT)
flags = [None] * w
dist.all_gather_object(flags, bool(ok))
if r == 0:
    print("RESULT", "OK" if all(flags) else "BAD", flags, flush=True)
dist.destroy_process_group()
'''


import pytest


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_forest_and_leaf_init_gloo(tmp_path, world):
    """world_size 2 and 3 (5 trees: uneven tree ownership) over gloo: the sharded forest + leaf-init protocol end to end on an
    oracle-backed engine."""
    script = tmp_path / "leaf_worker.py"
    script.write_text(LEAF_WORKER)
    env = dict(os.environ, REPO_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world, "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][0].split()
    assert line[1] == "OK", out.stdout[-2000:]


BUILD_WORKER = r'''
"""The WHOLE sharded build (ShardedBuilder.build: forest, leaf init with travelling heap rows, init_random, NN-descent with the
row-sharded candidate build (count-table all-gather + all-to-all-v of reverse-edge proposals), threshold all-gather, all-to-all of
half-updates, termination, final all-gather) over gloo, every engine primitive executed by the oracle on the OWN rows only.  The graph must be
identical to the single-process oracle build -- for any world size (the property tests/test_sharded_gpu.py checks on GPUs)."""
import os, sys
sys.path.insert(0, os.environ["REPO_ROOT"]); sys.path.insert(0, os.path.join(os.environ["REPO_ROOT"], "tests"))
import numpy as np, torch, torch.distributed as dist
dist.init_process_group("gloo")
from conftest import latent
from oracle import nnd_oracle as orc
from pynndescent_b200.sharded import ShardedBuilder
r, w = dist.get_rank(), dist.get_world_size()
N, D, K, LEAF, NT, T = 500, 10, 6, 16, 4, 8
X = latent(N, D, 4)
rs = np.random.RandomState(11)
rng_state = rs.randint(-2 ** 31 + 1, 2 ** 31 - 1, 3).astype(np.int64)
states = rs.randint(-2 ** 31 + 1, 2 ** 31 - 1, size=(NT, 3)).astype(np.int64)


def csr(la):
    sizes = (la >= 0).sum(1)
    offs = np.zeros(la.shape[0] + 1, np.int64); np.cumsum(sizes, out=offs[1:])
    return la[la >= 0].astype(np.int32), offs, int(sizes.max())


class OracleShardEngine:
    def __init__(self):
        self.N, self.k, self.device = N, K, torch.device("cpu")
        self.heap_idx = self.heap_dist = self.heap_flag = None
        self.lo, self.hi = 0, N
        self._ws = None
        self.stream = type("S", (), {"synchronize": lambda self: None})()
    def heap(self): return (self.heap_idx.numpy(), self.heap_dist.numpy(), self.heap_flag.numpy())
    def default_push_capacity(self): return 1 << 20
    def workspace(self, C_, T_, cap): return None
    def make_heap(self):
        a, b, c = orc.make_heap(N, K)
        self.heap_idx.copy_(torch.from_numpy(a)); self.heap_dist.copy_(torch.from_numpy(b)); self.heap_flag.copy_(torch.from_numpy(c))
    def forest(self, tree_states, leaf_size, angular, max_depth):
        la = orc.forest_leaf_array(X, np.ascontiguousarray(tree_states, np.int64), leaf_size, angular, max_depth, "level")
        pts, offs, ml = csr(la)
        return torch.from_numpy(pts), torch.from_numpy(offs), la.shape[0], ml
    def leaf_fused_range(self, pts, offs, l0, l1, max_leaf, pos0=None, npos=None, cap=None):
        if l1 == l0: return
        o = offs.numpy(); p = pts.numpy()
        la = np.full((l1 - l0, max_leaf), -1, np.int32)
        for i, l in enumerate(range(l0, l1)):
            m = p[o[l]: o[l + 1]]; la[i, : len(m)] = m
        orc.init_rp_tree(X, "euclidean", self.heap(), la)
    def heap_pack_rows(self, rows, n, oi, od, of, unpack):
        if n == 0: return
        idx = rows[:n].long()
        if not unpack:
            oi[:n] = self.heap_idx[idx]; od[:n] = self.heap_dist[idx]; of[:n] = self.heap_flag[idx]
        else:
            self.heap_idx[idx] = oi[:n]; self.heap_dist[idx] = od[:n]; self.heap_flag[idx] = of[:n]
    def init_random(self, rng): orc.init_random(X, "euclidean", self.heap(), rng)
    def set_row_range(self, lo, hi): self.lo, self.hi = lo, min(hi, N)
    # ---- the three steps of the row-sharded candidate build (csrc/k_cand.cu nnd_cand_*), restated on the oracle's rng and heap
    # push: this rank reads ONLY its own rows of the heap (the other rows are stale here, exactly like on the GPUs)
    def set_thresholds(self, thr): self.thr = None if thr is None else thr.numpy()
    def _tables(self, rng, T_):
        if getattr(self, "_tab", None) is None:
            self._tab = []
            for n in range(T_):
                st = (np.asarray(rng, np.int64) + n).copy()                    # utils.py:262 local_rng_state = rng_state + n
                self._tab.append(np.array([orc.tau_rand(st) for _ in range(N * K + 1)], np.float32))
        return self._tab
    def cand_local_counts(self, C_, T_, world, Bown):
        Bt = N // T_ + 1
        idx = self.heap()[0]
        cntR = np.zeros((T_, N), np.int64); valid = np.zeros(N, np.int64); dest = np.zeros(world, np.int64)
        for i in range(self.lo, self.hi):
            row = idx[i]; v = row >= 0
            valid[i] = v.sum()
            for n in range(T_):
                cntR[n, i] = ((row >= n * Bt) & (row < (n + 1) * Bt) & v).sum()
            for x in row[v]:
                dest[min(int(x) // Bown, world - 1)] += 1
        self.PR = np.concatenate([np.zeros((T_, 1), np.int64), np.cumsum(cntR, 1)], 1); self.PV = np.concatenate([[0], np.cumsum(valid)])
        out = np.zeros(5 * T_ + 1 + world, np.int64)
        for n in range(T_):
            s = min(n * Bt, N); e_ = min(s + Bt, N)
            out[5 * n: 5 * n + 5] = [self.PR[n, s], self.PV[s], self.PV[e_], self.PR[n, e_], self.PR[n, N]]
        out[5 * T_] = self.PV[N]; out[5 * T_ + 1:] = dest
        return out
    def cand_export(self, C_, rng, T_, tables_valid, world, Bown, glob, n_rec):
        Bt = N // T_ + 1
        tab = self._tables(rng, T_)
        idx, _, flg = self.heap()
        def draw(n, i, wv, wn):
            s = n * Bt; e_ = min(s + Bt, N); G = glob[6 * n: 6 * n + 6]
            prR = G[4] + self.PR[n, i]; prV = G[5] + self.PV[i]
            if i < s: return prR + wn
            if i < e_: return G[0] + (prV - G[1]) + wv
            return G[0] + (G[2] - G[1]) + (prR - G[3]) + wn
        self.fprio = np.zeros((N, K), np.float32)
        buckets = [[] for _ in range(world)]
        for i in range(self.lo, self.hi):
            ni = i // Bt; wv = 0
            for j in range(K):
                x = int(idx[i, j])
                if x < 0: continue
                nx = x // Bt
                pf = tab[ni][draw(ni, i, wv, 0)]
                self.fprio[i, j] = pf
                pr = pf
                if nx != ni:
                    wn = int(((idx[i, :j] >= nx * Bt) & (idx[i, :j] < (nx + 1) * Bt)).sum())
                    pr = tab[nx][draw(nx, i, 0, wn)]
                buckets[min(x // Bown, world - 1)].append((x, ((i * K + j) << 1) | int(flg[i, j] != 0), pr))
                wv += 1
        rec = np.zeros(max(n_rec, 1), np.dtype([("tgt", "<i4"), ("key", "<u4"), ("prio", "<f4")]))
        assert sum(len(b) for b in buckets) == n_rec
        q = 0
        for o in range(world):                                      # grouped by owner process in rank order
            for t in buckets[o][::-1]:                              # any order inside a group
                rec[q] = t; q += 1
        return torch.from_numpy(rec.view(np.uint8).copy())
    def cand_import_select(self, C_, T_, records, n):
        rec = records.numpy()[: n * 12].view(np.dtype([("tgt", "<i4"), ("key", "<u4"), ("prio", "<f4")]))
        idx, _, flg = self.heap()
        newc = np.full((N, C_), -1, np.int32); oldc = np.full((N, C_), -1, np.int32)
        by_t = {}
        for t, key, pr in rec:
            assert self.lo <= t < self.hi, "reverse-edge record delivered to the wrong owner"
            by_t.setdefault(int(t), []).append((int(key) >> 1, 1, np.float32(pr), int(key >> 1) // K, int(key) & 1))
        for v in range(self.lo, self.hi):
            ev = [(v * K + j, 0, self.fprio[v, j], int(idx[v, j]), int(flg[v, j] != 0)) for j in range(K) if idx[v, j] >= 0] + by_t.get(v, [])
            ev.sort(key=lambda t: (t[0], t[1]))                       # scan order; the forward edge first on a tie (self edge)
            hp = {1: (np.full(C_, np.inf, np.float32), newc[v]), 0: (np.full(C_, np.inf, np.float32), oldc[v])}
            for _, _, pr, cand, isnew in ev:
                orc.lib().orc_checked_heap_push(hp[isnew][0], hp[isnew][1], C_, float(pr), cand)
            for j in range(K):                                        # utils.py:311-318
                if idx[v, j] >= 0 and (newc[v] == idx[v, j]).any():
                    flg[v, j] = 0
        return newc, oldc
    def join_prepare(self, newc, oldc, cap):
        self.newc, self.oldc = newc, oldc
        return 1
    def join_emit_round(self, C_, cap, rnd):
        if rnd > 0: return 0
        thr = self.thr[:N]                                      # all-gathered from the owners (set_thresholds); local non-own rows are stale
        stride = 2 * C_
        recs = []
        for i in range(self.lo, self.hi):                      # own rows only; emission keys as in k_join / k_join_batch
            nw = self.newc[i][self.newc[i] >= 0]; ol = self.oldc[i][self.oldc[i] >= 0]
            mem = np.concatenate([nw, ol])
            for a in range(len(nw)):
                for b in range(a, len(mem)):
                    p, q = int(mem[a]), int(mem[b])
                    dd = np.float32(orc.dist("euclidean", X[p], X[q]))
                    base = (i << 20) | ((a * stride + b) << 1)
                    if dd < thr[p]: recs.append((p, q, dd, base))
                    if b != a and dd < thr[q]: recs.append((q, p, dd, base | 1))
        self.cur = recs
        return len(recs)
    def push_export(self, C_, cap, n, B, world):
        if n == 0: return torch.zeros(24, dtype=torch.uint8), [0] * world
        recs = sorted(self.cur, key=lambda t: min(t[0] // B, world - 1))          # grouped by owner, any order inside
        rec = np.zeros((n, 6), np.int32)
        rec[:, 0] = [t[0] for t in recs]; rec[:, 1] = [t[1] for t in recs]
        rec[:, 2] = np.array([t[2] for t in recs], np.float32).view(np.int32)
        seq = np.array([t[3] for t in recs], np.int64); rec[:, 4:6] = seq.view(np.int32).reshape(n, 2)
        counts = [int(sum(1 for t in recs if min(t[0] // B, world - 1) == o)) for o in range(world)]
        return torch.from_numpy(rec).view(torch.uint8).reshape(-1), counts
    def apply_records(self, C_, cap, buf, n):
        rec = buf[: n * 24].view(torch.int32).view(n, 6).numpy()
        seq = np.ascontiguousarray(rec[:, 4:6]).view(np.int64).ravel()
        hi_, hd_, hf_ = self.heap()
        c = 0
        for j in np.lexsort((seq, rec[:, 0])):                  # per target row in ascending emission sequence
            t = int(rec[j, 0]); assert self.lo <= t < self.hi
            c += orc.lib().orc_checked_flagged_heap_push(hd_[t], hi_[t], hf_[t], K, float(rec[j, 2:3].view(np.float32)[0]), int(rec[j, 1]), 1)
        return c
    def deheap_sort(self, correction):
        a, b, _ = self.heap()
        orc.deheap_sort(a, b)
        return self.heap_idx, torch.from_numpy(np.sqrt(b) if correction == 1 else b.copy())


NO_TREES = os.environ.get("NND_TEST_NO_TREES") == "1"      # tree_init=False: every row is filled by init_random (the all-gather branch)
e = OracleShardEngine()
sb = ShardedBuilder(e)
rng_run = rng_state.copy()
idx_t, dist_t, iters, c_hist = sb.build(rng_run, None if NO_TREES else states, LEAF, False, K, T, 8, 0.001, 1)
# single-process oracle with the same seeds (forest scheme "level" from the same tree states)
ref = orc.make_heap(N, K)
if not NO_TREES:
    la = orc.forest_leaf_array(X, states, LEAF, False, 200, "level")
    orc.init_rp_tree(X, "euclidean", ref, la)
rng2 = rng_state.copy()
orc.init_random(X, "euclidean", ref, rng2)
assert np.array_equal(rng_run, rng2), "init_random must advance the rng identically on every rank"
ch = np.zeros(8, np.int64)
it2 = orc.lib().orc_nn_descent_internal(X, N, D, 0, ref[0], ref[1], ref[2], K, K, rng2, T, 8, 0.001, 16384, ch)
orc.deheap_sort(ref[0], ref[1])
ok = (np.array_equal(idx_t.numpy(), ref[0]) and np.array_equal(dist_t.numpy(), np.sqrt(ref[1])) and iters == it2
      and list(c_hist) == [int(c) for c in ch[:it2]])
flags = [None] * w
dist.all_gather_object(flags, bool(ok))
if r == 0:
    print("RESULT", "OK" if all(flags) else "BAD", flags, iters, it2, list(c_hist), [int(c) for c in ch[:it2]], flush=True)
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world,no_trees", [(2, False), (3, False), (2, True)])
def test_whole_sharded_build_equals_single_process_oracle_gloo(tmp_path, world, no_trees):
    script = tmp_path / "build_worker.py"
    script.write_text(BUILD_WORKER)
    env = dict(os.environ, REPO_ROOT=ROOT, MASTER_ADDR="127.0.0.1", NND_TEST_NO_TREES="1" if no_trees else "0")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world, "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][0].split()
    assert line[1] == "OK", out.stdout[-2000:]
