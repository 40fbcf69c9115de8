"""Checksum fixture for the FULL-SIZE parity test (BASELINE.json configs[1]: 1M x 128 float32 euclidean k=15).

Runs the CPU oracle (forest scheme "level" = the GPU forest's, emulated numba threads T=8) on the bench workload
(latent-16 generator, seed 0, random_state 42) -- about a minute on 8 cores -- and stores sha256(indices), sha256(distances),
the per-iteration accepted-push counts and a few literal rows in tests/golden/cfg2_full_oracle.json.  The GPU test
(tests/test_zfull_size_gpu.py) must reproduce all of them: a 1,000,000 x 15 graph, bit for bit.

    python tests/golden/make_full_size_checksum.py [cfg2|cfg3|cfg5]
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
           "cfg3": (2_000_000, 256, 15, "cosine"),           # BASELINE.json configs[2] (several minutes of oracle time)
           "cfg5": (1_000_000, 784, 30, "correlation")}      # BASELINE.json configs[4] build (a quarter of an hour)

if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    n, d, k, metric = CONFIGS[name]
    orc.lib().orc_set_num_threads(os.cpu_count())
    X = latent(n, d, 0)
    t0 = time.time()
    o = orc.OracleNNDescent(X, metric, n_neighbors=k, random_state=42, n_threads=8, forest="level")
    gi, gd = o.neighbor_graph
    out = {"n": n, "d": d, "k": k, "metric": metric, "data": "conftest.latent(n, d, 0)", "random_state": 42, "T": 8,
           "forest": "level", "iters": int(o.n_iters_run), "c_hist": [int(c) for c in o.c_hist[: o.n_iters_run]],
           "idx_sha256": hashlib.sha256(np.ascontiguousarray(gi, np.int32).tobytes()).hexdigest(),
           "dist_dtype": str(np.asarray(gd).dtype),
           "dist_sha256": hashlib.sha256(np.ascontiguousarray(gd).tobytes()).hexdigest(),
           "rows": {str(r): {"idx": gi[r].tolist(), "dist": [float(v) for v in gd[r]]} for r in (0, 1, n // 2 - 1, n - 1)},
           "oracle_seconds": round(time.time() - t0, 1)}
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", name + "_full_oracle.json"), "w"), indent=1)
    print(json.dumps({kk: out[kk] for kk in ("iters", "c_hist", "idx_sha256", "dist_sha256", "oracle_seconds")}))
