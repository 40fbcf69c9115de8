"""CPU checks (oracle-driven) of the invariants the schedule switches of libnnd_b200 rely on.

NND_CAND_SKIP (default on): a vertex has new candidates iff it is an endpoint of a new-flagged edge, and the candidate lists of
vertices WITHOUT new candidates are never read by the local join -- blanking them must leave the whole build unchanged.
The replay filters of k_apply (a source already in the target's row / a repeated (target, source)): dropping those half-updates
before the ordered replay leaves heaps and accepted-push counts unchanged.
"""
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT, latent


def _build(X, k, blank_inactive):
    from sklearn.utils import check_random_state
    from oracle import nnd_oracle as orc
    n = X.shape[0]
    rs = check_random_state(5)
    rng_state = rs.randint(orc.INT32_MIN, orc.INT32_MAX, 3).astype(np.int64)
    leaf_array, _ = orc.make_forest(X, k, 4, None, rs, False, 200, "level")
    heap = orc.make_heap(n, k)
    orc.init_rp_tree(X, "euclidean", heap, leaf_array)
    orc.init_random(X, "euclidean", heap, rng_state)
    c_hist, inactive = [], []
    for it in range(12):
        flags_before = heap[2].copy(); idx_before = heap[0].copy()
        newc, oldc = orc.new_build_candidates(heap, k, rng_state, 8)
        has_new = (newc >= 0).any(1)
        # the claim behind the switch: has_new == "touches a new-flagged edge"
        touches = np.zeros(n, bool)
        new_edge = (flags_before != 0) & (idx_before >= 0)
        touches[new_edge.any(1)] = True
        touches[idx_before[new_edge]] = True
        assert np.array_equal(has_new, touches)
        if blank_inactive:
            oldc[~has_new] = -1
        inactive.append(int((~has_new).sum()))
        c, _ = orc.process_candidates(X, "euclidean", heap, newc, oldc, 8)
        c_hist.append(c)
        if c <= 0.001 * k * n:
            break
    return heap, c_hist, inactive


def test_candidates_of_vertices_without_new_candidates_are_dead():
    X = latent(6000, 24, 11)
    h0, c0, _ = _build(X, 10, False)
    h1, c1, inactive = _build(X, 10, True)
    assert c0 == c1
    for a, b in zip(h0, h1):
        assert np.array_equal(a, b)
    assert inactive[-1] > 0.5 * X.shape[0]            # the late iterations are where the switch pays


def test_present_and_repeated_half_updates_are_dead():
    """k_apply drops, before the ordered replay, every half-update (t <- o, d) whose o already sits in row t as loaded and all but the
    first (in emission order) of several with the same (t, o): neither can be accepted (while o is in the row the duplicate test
    rejects it, once evicted the root is at most d), so heaps, flags and the accepted-push count must not change.  Replayed here with
    the oracle's own push on three iterations of a build, with and without the filter."""
    from sklearn.utils import check_random_state
    from oracle import nnd_oracle as orc
    X = latent(3000, 16, 21); n, k = X.shape[0], 10
    rs = check_random_state(9)
    rng_state = rs.randint(orc.INT32_MIN, orc.INT32_MAX, 3).astype(np.int64)
    la, _ = orc.make_forest(X, k, 3, None, rs, False, 200, "level")
    heap = orc.make_heap(n, k)
    orc.init_rp_tree(X, "euclidean", heap, la); orc.init_random(X, "euclidean", heap, rng_state)
    dropped = 0
    for it in range(3):
        newc, oldc = orc.new_build_candidates(heap, k, rng_state, 8)
        thr = heap[1][:, 0].copy()
        p, q, dd = orc.generate_updates(X, "euclidean", newc, oldc, thr)
        # half-updates in emission order: (p <- q), then (q <- p) when p != q
        recs = []
        for a, b, d_ in zip(p.tolist(), q.tolist(), dd.tolist()):
            recs.append((a, b, np.float32(d_)))
            if a != b:
                recs.append((b, a, np.float32(d_)))
        results = []
        for filt in (False, True):
            h = tuple(x.copy() for x in heap)
            start_rows = h[0].copy()
            seen, c = set(), 0
            for t, o, d_ in recs:
                if filt:
                    if (t, o) in seen or (start_rows[t] == o).any():
                        dropped += 1
                        continue
                    seen.add((t, o))
                c += orc.lib().orc_checked_flagged_heap_push(h[1][t], h[0][t], h[2][t], k, float(d_), int(o), 1)
            results.append((h, c))
        (h0, c0), (h1, c1) = results
        assert c0 == c1 and all(np.array_equal(a, b) for a, b in zip(h0, h1))
        for dst, src in zip(heap, h0):
            dst[...] = src
    assert dropped > 1000


def test_ab_harness_switch_names_exist():
    """scripts/r2_ab.py A/Bs the schedule switches by environment variable: a misspelt name would silently compare the default
    with itself.  Every name it sets must be read by the library (csrc) or the host package."""
    import importlib.util
    import re
    spec = importlib.util.spec_from_file_location("r2_ab", os.path.join(ROOT, "scripts", "r2_ab.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    names = {k for cfg in mod.CONFIGS.values() for k in cfg}
    assert names and all(n.startswith("NND_") for n in names)
    src = ""
    for base in ("pynndescent_b200/csrc", "pynndescent_b200"):
        d = os.path.join(ROOT, base)
        for f in os.listdir(d):
            if f.endswith((".cu", ".cuh", ".py")):
                src += open(os.path.join(d, f)).read()
    read = set(re.findall(r'"(NND_[A-Z0-9_]+)"', src))
    assert names <= read, sorted(names - read)
