"""CPU checks (oracle-driven) of the invariants the default-OFF schedule switches of libnnd_b200 rely on.

NND_CAND_SKIP: a vertex has new candidates iff it is an endpoint of a new-flagged edge, and the candidate lists of
vertices WITHOUT new candidates are never read by the local join -- blanking them must leave the whole build unchanged.
NND_JOIN_ORDER / batching: covered on the GPU by the any-blocking invariance tests; the protocol of k_join_pipe by
scripts/pipe_protocol_sim.py (run here so that it stays in the CPU suite).
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


def test_pipe_protocol_simulation():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "pipe_protocol_sim.py"), "400"], capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr


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
