"""Golden fixtures for the query graph walk (seam 3), produced by the REAL reference in the build container:
the prepared search structures of a small index + queries + the reference's answers.
    python tests/golden/make_golden_query.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_shim import import_reference  # noqa: E402

os.environ["NUMBA_NUM_THREADS"] = "8"
pynndescent = import_reference()
import numba  # noqa: E402

numba.set_num_threads(8)
OUT = os.path.dirname(os.path.abspath(__file__))


def latent(n, d, seed=0, latent_dim=16):
    rs = np.random.RandomState(seed)
    Z = rs.standard_normal((n, latent_dim)).astype(np.float32)
    A = rs.standard_normal((latent_dim, d)).astype(np.float32) / np.sqrt(latent_dim)
    return (Z @ A + 0.05 * rs.standard_normal((n, d)).astype(np.float32)).astype(np.float32)


def dump(name, metric, n, d, k_index, seed, nq, k_query, eps, data_fn=None):
    both = (data_fn or latent)(n + nq, d, seed)
    X, Q = both[:n], both[n:]
    idx = pynndescent.NNDescent(X, metric, n_neighbors=k_index, random_state=seed)
    ng_idx, ng_dist = idx._neighbor_graph[0].copy(), idx._neighbor_graph[1].copy()   # raw (alternative) distances
    rng_before_prepare = idx.rng_state.copy()
    idx.prepare()
    tree = idx._search_forest[0]
    out = dict(
        X=X, Q=Q, metric=metric, k_index=k_index, k_query=k_query, epsilon=eps,
        data=idx._raw_data, indptr=idx._search_graph.indptr.astype(np.int32), indices=idx._search_graph.indices.astype(np.int32),
        hyperplanes=np.asarray(tree.hyperplanes, np.float32), offsets=np.asarray(tree.offsets, np.float32),
        children=np.asarray(tree.children, np.int32), tree_indices=np.asarray(tree.indices, np.int32),
        vertex_order=np.asarray(idx._vertex_order, np.int32), min_distance=np.float32(idx._min_distance),
        search_rng_state=idx.search_rng_state.copy(), n_neighbors=k_index, ng_idx=ng_idx, ng_dist=ng_dist,
        rng_before_prepare=rng_before_prepare,
    )
    Qc = np.asarray(Q).astype(np.float32, order="C")
    hi, hd, _ = idx._search_function(Qc, k_query, eps, idx._visited, idx.search_rng_state)
    out["heap_idx"], out["heap_dist"] = hi.copy(), hd.copy()
    qi, qd = idx.query(Q, k=k_query, epsilon=eps)
    out["out_idx"], out["out_dist"] = qi, qd
    # a second epsilon / k for coverage
    qi2, qd2 = idx.query(Q, k=5, epsilon=0.3)
    out["out_idx_k5_e03"], out["out_dist_k5_e03"] = qi2, qd2
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    true = np.argsort(((Q[:, None, :] - X[None, :, :]) ** 2).sum(-1), 1)[:, :k_query] if metric == "euclidean" else None
    if true is not None:
        print(name, "query recall", np.mean([len(np.intersect1d(true[i], qi[i])) / k_query for i in range(nq)]))
    print(name, "saved: graph nnz", out["indices"].shape[0], "tree nodes", out["children"].shape[0])


def binary(n, d, seed):
    return (latent(n, d, seed) > 0.35).astype(np.float32)


def nonneg(n, d, seed):
    X = np.abs(latent(n, d, seed)).astype(np.float32)
    X[X < 0.15] = 0.0
    return X


if __name__ == "__main__":
    which = sys.argv[1:] or ["base", "metrics"]
    if "base" in which:
        dump("query_euclid_3k32", "euclidean", 3000, 32, 15, 42, 200, 10, 0.1)
        dump("query_cosine_2k24", "cosine", 2000, 24, 10, 7, 150, 10, 0.1)
    if "metrics" in which:      # SURVEY 8f-2 metrics through prepare() and query()
        dump("query_manhattan_2k20", "manhattan", 2000, 20, 12, 11, 150, 10, 0.1)
        dump("query_jaccard_2k40", "jaccard", 2000, 40, 10, 12, 150, 10, 0.1, data_fn=binary)
        dump("query_hellinger_2k24", "hellinger", 2000, 24, 10, 13, 150, 10, 0.1, data_fn=nonneg)
