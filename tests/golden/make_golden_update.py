"""Golden fixtures for NNDescent.update() (pynndescent_.py:2381-2553) from the REAL reference.

Run in the build container only:   python tests/golden/make_golden_update.py
Writes tests/golden/update_*.npz.  numba threads pinned to T=8.  For every case the reference index is built on
`xs_orig`, then `update(xs_fresh, xs_updated, updated_indices)` is called; rptree_leaf_array is wrapped so that the
leaf_array of BOTH forests (the constructor's and the update's) is captured -- with those injected, everything
downstream (init_from_neighbor_graph -> init_rp_tree -> nn_descent) is deterministic and must be reproduced exactly.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_shim import import_reference  # noqa: E402

T = 8
os.environ["NUMBA_NUM_THREADS"] = str(T)
pynndescent = import_reference()
import numba  # noqa: E402
from pynndescent import pynndescent_ as pp  # noqa: E402

numba.set_num_threads(T)
OUT = os.path.dirname(os.path.abspath(__file__))


def latent(n, d, seed=0, latent_dim=16):
    rs = np.random.RandomState(seed)
    Z = rs.standard_normal((n, latent_dim)).astype(np.float32)
    A = rs.standard_normal((latent_dim, d)).astype(np.float32) / np.sqrt(latent_dim)
    return (Z @ A + 0.05 * rs.standard_normal((n, d)).astype(np.float32)).astype(np.float32)


captured = []
_orig_leaf_array = pp.rptree_leaf_array


def _capture(forest):
    la = _orig_leaf_array(forest)
    captured.append(np.array(la, np.int32))
    return la


pp.rptree_leaf_array = _capture


def case(name, metric, n0, n_fresh, n_upd, d, k, seed, prepare_first=False, second_update=False):
    X = latent(n0 + n_fresh + n_upd, d, seed)
    xs_orig, xs_fresh, xs_upd = X[:n0], X[n0:n0 + n_fresh], X[n0 + n_fresh:]
    rs = np.random.RandomState(seed + 1)
    upd_idx = np.sort(rs.choice(n0, n_upd, replace=False)).astype(np.int64) if n_upd else None
    captured.clear()
    idx = pynndescent.NNDescent(xs_orig.copy(), metric=metric, n_neighbors=k, random_state=seed, n_jobs=T)
    out = {"xs_orig": xs_orig, "xs_fresh": xs_fresh, "xs_updated": xs_upd, "k": k, "seed": seed, "T": T,
           "updated_indices": upd_idx if upd_idx is not None else np.zeros(0, np.int64),
           "leaf_array_build": captured[0], "n_trees_build": idx.n_trees,
           "graph0_idx": idx._neighbor_graph[0].copy(), "graph0_raw_dist": idx._neighbor_graph[1].copy(),
           "rng_state_after_build": idx.rng_state.copy()}
    if prepare_first:
        idx.prepare()
    idx.update(xs_fresh=xs_fresh if n_fresh else None, xs_updated=xs_upd if n_upd else None, updated_indices=upd_idx)
    out["leaf_array_update"] = captured[1]
    out["n_trees_after"] = idx.n_trees
    out["graph1_idx"], out["graph1_raw_dist"] = idx._neighbor_graph[0].copy(), idx._neighbor_graph[1].copy()
    gi, gd = idx.neighbor_graph
    out["graph1_dist"] = np.asarray(gd)
    out["rng_state_after_update"] = idx.rng_state.copy()
    if second_update:                      # tree count must stay at n_trees_after_update (reference test :646-653)
        extra = latent(60, d, seed + 7)
        idx.update(xs_fresh=extra)
        out["xs_fresh2"] = extra
        out["leaf_array_update2"] = captured[2]
        out["graph2_idx"], out["graph2_raw_dist"] = idx._neighbor_graph[0].copy(), idx._neighbor_graph[1].copy()
        out["n_trees_after2"] = idx.n_trees
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "ok: trees", out["n_trees_build"], "->", out["n_trees_after"], "graph", out["graph1_idx"].shape)


if __name__ == "__main__":
    case("update_fresh_euclid_1200x16", "euclidean", 1000, 200, 0, 16, 10, 11, second_update=True)
    case("update_changed_euclid_1200x16", "euclidean", 1100, 100, 60, 16, 10, 12)
    case("update_changed_cosine_1000x20", "cosine", 900, 100, 40, 20, 8, 13)
    case("update_only_changed_euclid_800x12", "euclidean", 800, 0, 50, 12, 10, 14)
