"""Checksum + recall fixtures for the FULL-SIZE parity tests and bench lines (BASELINE.json configs[1..4]).

Runs the CPU oracle on a bench workload (latent-16 generator, seed 0, random_state 42, emulated numba threads T=8) twice:

* forest scheme "level" (= the GPU forest's): sha256(indices), sha256(distances), per-iteration accepted-push counts and a
  few literal rows -- the GPU build must reproduce all of them bit for bit (tests/test_zfull_size_gpu.py, bench.py's
  `parity` field);
* forest scheme "dfs" (= the reference's own forest, pynndescent/rp_trees.py:2177-2302): its recall@k against exact
  brute force on the SAME fixed 1000-row sample (`np.random.RandomState(0).choice(n, 1000, replace=False)`, the sample of
  bench.recall_at_k) -- the number `north_star`'s "recall within 0.5 pt of the reference on the same seed" is checked
  against.  The exact neighbours of the sample rows are stored too, so that the GPU side needs no brute force of its own.

    python tests/golden/make_full_size_checksum.py [cfg2|cfg3|cfg4|cfg5] [--level] [--dfs]     (default: both)

Oracle time on 8 cores: cfg2 about a minute per scheme, cfg3 several minutes, cfg4 (10M x 128) and cfg5 a quarter of an hour.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import latent  # noqa: E402
from oracle import nnd_oracle as orc  # noqa: E402

CONFIGS = {"cfg2": (1_000_000, 128, 15, "euclidean"),       # BASELINE.json configs[1]
           "cfg3": (2_000_000, 256, 15, "cosine"),           # BASELINE.json configs[2]
           "cfg4": (10_000_000, 128, 15, "euclidean"),       # BASELINE.json configs[3] = the metric's own configuration
           "cfg5": (1_000_000, 784, 30, "correlation"),      # BASELINE.json configs[4] build
           "tiny": (20_000, 32, 10, "euclidean"), "tinycos": (20_000, 32, 10, "cosine")}   # self-test of this script


def exact_knn(X, rows, k, metric, block=500_000):
    """Exact k nearest neighbours of X[rows] in X: float32 GEMM shortlist of 4k per block, float64 re-rank."""
    Xn = X
    if metric == "correlation":
        Xn = X - X.mean(1, keepdims=True)
    if metric in ("cosine", "correlation"):
        Xn = Xn / np.maximum(np.linalg.norm(Xn, axis=1, keepdims=True), 1e-30)
    Q = Xn[rows]
    short_i, short_d = [], []
    for c0 in range(0, X.shape[0], block):
        B = Xn[c0:c0 + block]
        if metric in ("cosine", "correlation"):
            dm = 1.0 - Q @ B.T
        else:
            dm = (B * B).sum(1)[None, :] - 2.0 * (Q @ B.T)
        m = min(4 * k, dm.shape[1])
        part = np.argpartition(dm, m - 1, axis=1)[:, :m]
        short_i.append(part + c0); short_d.append(np.take_along_axis(dm, part, 1))
    ci = np.concatenate(short_i, 1); cd = np.concatenate(short_d, 1)
    m = min(4 * k, ci.shape[1])
    sel = np.argpartition(cd, m - 1, axis=1)[:, :m]
    ci = np.take_along_axis(ci, sel, 1)
    out = np.empty((len(rows), k), np.int64)
    for i in range(len(rows)):
        a = Q[i].astype(np.float64); Bc = Xn[ci[i]].astype(np.float64)
        dd = 1.0 - Bc @ a if metric in ("cosine", "correlation") else ((Bc - a) ** 2).sum(1)
        out[i] = ci[i][np.argsort(dd, kind="stable")[:k]]
    return out


def recall(true, gi, rows):
    return float(np.mean([len(np.intersect1d(true[i], gi[r])) for i, r in enumerate(rows)]) / true.shape[1])


if __name__ == "__main__":
    name = next((a for a in sys.argv[1:] if not a.startswith("--")), "cfg2")
    do_level = "--level" in sys.argv or "--dfs" not in sys.argv
    do_dfs = "--dfs" in sys.argv or "--level" not in sys.argv
    n, d, k, metric = CONFIGS[name]
    path = os.path.join(ROOT, "tests", "golden", name + "_full_oracle.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    orc.lib().orc_set_num_threads(os.cpu_count())
    X = latent(n, d, 0)
    rows = np.random.RandomState(0).choice(n, 1000, replace=False)
    if "exact_knn_of_sample" not in out:
        t0 = time.time()
        true = exact_knn(X, rows, k, metric)
        out["recall_sample"] = "np.random.RandomState(0).choice(n, 1000, replace=False)  (= bench.recall_at_k's rows)"
        out["exact_knn_of_sample"] = true.tolist()
        print("exact kNN of the sample: %.1f s" % (time.time() - t0), flush=True)
    true = np.asarray(out["exact_knn_of_sample"])
    if do_level:
        t0 = time.time()
        o = orc.OracleNNDescent(X, metric, n_neighbors=k, random_state=42, n_threads=8, forest="level")
        gi, gd = o.neighbor_graph
        out.update({"n": n, "d": d, "k": k, "metric": metric, "data": "conftest.latent(n, d, 0)", "random_state": 42, "T": 8,
                    "forest": "level", "iters": int(o.n_iters_run), "c_hist": [int(c) for c in o.c_hist[: o.n_iters_run]],
                    "idx_sha256": hashlib.sha256(np.ascontiguousarray(gi, np.int32).tobytes()).hexdigest(),
                    "dist_dtype": str(np.asarray(gd).dtype),
                    "dist_sha256": hashlib.sha256(np.ascontiguousarray(gd).tobytes()).hexdigest(),
                    "rows": {str(r): {"idx": gi[r].tolist(), "dist": [float(v) for v in gd[r]]} for r in (0, 1, n // 2 - 1, n - 1)},
                    "recall_level_oracle": recall(true, gi, rows),
                    "oracle_seconds": round(time.time() - t0, 1)})
        json.dump(out, open(path, "w"), indent=1)
        print(json.dumps({kk: out[kk] for kk in ("iters", "c_hist", "idx_sha256", "dist_sha256", "recall_level_oracle", "oracle_seconds")}), flush=True)
        del o, gi, gd
    if do_dfs:
        t0 = time.time()
        o = orc.OracleNNDescent(X, metric, n_neighbors=k, random_state=42, n_threads=8, forest="dfs")
        gi, gd = o.neighbor_graph
        out.update({"recall_dfs_oracle": recall(true, gi, rows), "dfs_iters": int(o.n_iters_run),
                    "dfs_c_hist": [int(c) for c in o.c_hist[: o.n_iters_run]], "dfs_oracle_seconds": round(time.time() - t0, 1)})
        json.dump(out, open(path, "w"), indent=1)
        print(json.dumps({kk: out[kk] for kk in ("recall_dfs_oracle", "dfs_iters", "dfs_oracle_seconds")}), flush=True)
