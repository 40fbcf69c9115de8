"""Generate golden fixtures from the REAL reference (lmcinnes/pynndescent v0.6.0).

Run in the build container only:   python tests/golden/make_golden.py
Writes tests/golden/*.npz.  numba threads are pinned to T=8 (the reference is
deterministic only for a fixed (seed, T)).
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
from sklearn.utils import check_random_state  # noqa: E402
from pynndescent import pynndescent_ as pp  # noqa: E402
from pynndescent import utils as pu, rp_trees as rpt, distances as pd  # noqa: E402

numba.set_num_threads(T)
OUT = os.path.dirname(os.path.abspath(__file__))
INT32_MIN = np.iinfo(np.int32).min + 1
INT32_MAX = np.iinfo(np.int32).max - 1


def latent(n, d, seed=0, latent_dim=16):
    rs = np.random.RandomState(seed)
    Z = rs.standard_normal((n, latent_dim)).astype(np.float32)
    A = rs.standard_normal((latent_dim, d)).astype(np.float32) / np.sqrt(latent_dim)
    return (Z @ A + 0.05 * rs.standard_normal((n, d)).astype(np.float32)).astype(np.float32)


def kat():
    out = {}
    for name, st in (("a", [-5, 123456789, -987654321]), ("b", [2147483646, -2147483647, 42]), ("c", [1, 2, 3])):
        s = np.array(st, np.int64)
        out[f"tau_{name}_state0"] = s.copy()
        out[f"tau_{name}_ints"] = np.array([pu.tau_rand_int(s) for _ in range(64)], np.int32)
        out[f"tau_{name}_state1"] = s.copy()
        s = np.array(st, np.int64)
        out[f"tau_{name}_floats"] = np.array([pu.tau_rand(s) for _ in range(64)], np.float32)
    # heap push KAT + a random longer one
    rs = np.random.RandomState(7)
    for name, k, n in (("h5", 5, 40), ("h15", 15, 200), ("h30", 30, 400)):
        pri = np.full(k, np.inf, np.float32); ind = np.full(k, -1, np.int32); flg = np.zeros(k, np.uint8)
        d = np.round(rs.random_sample(n), 2).astype(np.float32)  # rounding -> ties
        q = rs.randint(0, 3 * k, n).astype(np.int32)
        f = rs.randint(0, 2, n).astype(np.uint8)
        ret = np.array([pu.checked_flagged_heap_push(pri, ind, flg, d[i], q[i], f[i]) for i in range(n)], np.int32)
        out[f"{name}_d"], out[f"{name}_q"], out[f"{name}_f"], out[f"{name}_ret"] = d, q, f, ret
        out[f"{name}_pri"], out[f"{name}_ind"], out[f"{name}_flg"] = pri.copy(), ind.copy(), flg.copy()
        si, sd = pu.deheap_sort(ind.copy()[None, :], pri.copy()[None, :])
        out[f"{name}_sorted_ind"], out[f"{name}_sorted_pri"] = si[0], sd[0]
        # simple (unchecked) push as used by the query walk
        pri2 = np.full(k, np.inf, np.float32); ind2 = np.full(k, -1, np.int32)
        ret2 = np.array([pu.simple_heap_push(pri2, ind2, d[i], q[i]) for i in range(n)], np.int32)
        out[f"{name}_simple_ret"], out[f"{name}_simple_pri"], out[f"{name}_simple_ind"] = ret2, pri2, ind2
    # distances
    rs = np.random.RandomState(3)
    A = rs.standard_normal((24, 37)).astype(np.float32)
    A[5] = 0.0
    A[6] = A[7]
    A[8] = 2.5  # zero variance row
    out["dist_A"] = A
    fns = {"sqeuclidean": pd.squared_euclidean, "cosine": pd.alternative_cosine, "dot": pd.alternative_dot,
           "correlation": pd.correlation}
    for name, fn in fns.items():
        out[f"dist_{name}"] = np.array([[fn(A[i], A[j]) for j in range(24)] for i in range(24)], np.float32)
    out["corr_cosine"] = pd.correct_alternative_cosine(out["dist_cosine"])
    np.savez_compressed(os.path.join(OUT, "kat.npz"), **out)


def staged(name, X, metric, k, seed, n_trees=None, leaf_size=None, n_iters=None, metric_kwds=None):
    """Replay NNDescent.__init__ (pynndescent_.py:976-1269) stage by stage with the reference's
    own numba functions and dump every intermediate."""
    out = {"X": X, "k": k, "seed": seed, "T": T}
    n = X.shape[0]
    if n_trees is None:
        n_trees = max(3, min(12, int(round(2.0 * np.log10(n)))))
    if n_iters is None:
        n_iters = max(5, int(round(np.log2(n))))
    idx = pynndescent.NNDescent(X, metric, metric_kwds=metric_kwds, n_neighbors=k, random_state=seed, n_trees=n_trees,
                                leaf_size=leaf_size, n_iters=n_iters)
    gi, gd = idx.neighbor_graph
    out["graph_idx"], out["graph_dist"] = gi, gd
    dist = idx._distance_func
    data = idx._raw_data
    angular = idx._angular_trees
    rs = check_random_state(seed)
    rng_state = rs.randint(INT32_MIN, INT32_MAX, 3).astype(np.int64)
    _search = rs.randint(INT32_MIN, INT32_MAX, 3).astype(np.int64)
    out["rng_state0"] = rng_state.copy()
    forest = rpt.make_forest(data, k, n_trees, leaf_size, rng_state, rs, None, angular, False, max_depth=200)
    leaf_array = rpt.rptree_leaf_array(forest)
    out["leaf_array"] = leaf_array
    rs2 = check_random_state(seed); rs2.randint(INT32_MIN, INT32_MAX, 3); rs2.randint(INT32_MIN, INT32_MAX, 3)
    out["tree_rng_states"] = rs2.randint(INT32_MIN, INT32_MAX, size=(n_trees, 3)).astype(np.int64)
    C = min(60, k)
    heap = pu.make_heap(n, k)
    pp.init_rp_tree(data, dist, heap, leaf_array)
    out["leaf_heap_idx"], out["leaf_heap_dist"], out["leaf_heap_flag"] = [a.copy() for a in heap]
    pp.init_random(k, data, heap, dist, rng_state)
    out["init_heap_idx"], out["init_heap_dist"], out["init_heap_flag"] = [a.copy() for a in heap]
    out["rng_state1"] = rng_state.copy()
    # iterations, dumping the first two in detail
    cs = []
    block = 16384
    n_blocks = n // block
    nthr = numba.get_num_threads()
    maxu = int((C ** 2 + C * (C - 1) / 2) * block / nthr) + 1024
    ua = np.empty((nthr, maxu, 3), np.float32); nu = np.zeros(nthr, np.int32)
    for it in range(n_iters):
        newc, oldc = pu.new_build_candidates(heap, C, rng_state, nthr)
        if it < 2:
            out[f"it{it}_new"], out[f"it{it}_old"], out[f"it{it}_flag_after_cand"] = newc.copy(), oldc.copy(), heap[2].copy()
            out[f"it{it}_thr"] = heap[1][:, 0].copy()
        c = pp.process_candidates(data, dist, heap, newc, oldc, n_blocks, block, nthr, ua, nu)
        cs.append(c)
        if it < 2:
            out[f"it{it}_heap_idx"], out[f"it{it}_heap_dist"], out[f"it{it}_heap_flag"] = [a.copy() for a in heap]
        if c <= 0.001 * k * n:
            break
    out["c_hist"] = np.array(cs, np.int64)
    out["final_heap_idx"], out["final_heap_dist"] = heap[0].copy(), heap[1].copy()
    si, sd = pu.deheap_sort(heap[0].copy(), heap[1].copy())
    assert np.array_equal(si, idx._neighbor_graph[0]), "staged replay diverged from NNDescent()"
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
    print(name, "ok: iters", len(cs), "c", cs)


if __name__ == "__main__":
    kat()
    staged("cfg1_euclid_1k16", np.random.RandomState(0).standard_normal((1000, 16)).astype(np.float32),
           "euclidean", 10, 42)
    staged("latent_euclid_3k32", latent(3000, 32, 1), "euclidean", 15, 42)
    staged("latent_cosine_2k24", latent(2000, 24, 2), "cosine", 10, 7)
    staged("latent_corr_1500x20", latent(1500, 20, 3), "correlation", 12, 11)
    staged("latent_dot_1500x20", np.abs(latent(1500, 20, 4)), "dot", 8, 5)
