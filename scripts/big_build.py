"""One full-size build with stage timings, graph checksum and recall (dev tool; GPU box).

    python scripts/big_build.py N d k [metric] [reps] [--json=path]

Prints per-rep stage timings / c_hist and, for the last graph, sha256 of indices and distances and recall@k on the
fixed 1000-row sample bench.py uses (same rows as tests/golden/*_full_oracle.json).
"""
import hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import latent, recall_at_k
from pynndescent_b200 import NNDescent

args = [a for a in sys.argv[1:] if not a.startswith("--")]
opts = dict(a[2:].split("=", 1) for a in sys.argv[1:] if a.startswith("--") and "=" in a)
N, d, k = int(args[0]), int(args[1]), int(args[2])
metric = args[3] if len(args) > 3 else "euclidean"
reps = int(args[4]) if len(args) > 4 else 2
print("cpu_count", os.cpu_count(), "gpu", torch.cuda.get_device_name(0), flush=True)
t0 = time.perf_counter()
X = latent(N, d)
print(f"data {time.perf_counter() - t0:.1f}s", flush=True)
out = {"n": N, "d": d, "k": k, "metric": metric, "reps": []}
for rep in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    idx = NNDescent(X, metric, n_neighbors=k, random_state=42)
    torch.cuda.synchronize(); t = time.perf_counter() - t0
    r = {"total_s": t, "edges_per_s": N * k / t, "timings_ms": {a: round(b * 1e3, 1) for a, b in idx.timings.items()},
         "iters": idx.n_iters_run, "c_hist": [int(c) for c in idx.c_hist], "join_stats": [int(v) for v in idx.join_stats],
         "mem_gb": torch.cuda.max_memory_allocated() / 1e9}
    out["reps"].append(r)
    print(json.dumps(r), flush=True)
    if rep < reps - 1:
        del idx
gi, gd = idx.neighbor_graph
out["idx_sha256"] = hashlib.sha256(np.ascontiguousarray(gi).tobytes()).hexdigest()
out["dist_sha256"] = hashlib.sha256(np.ascontiguousarray(gd).tobytes()).hexdigest()
out["recall_at_k"] = recall_at_k(X, gi, k, metric)
print(json.dumps({kk: out[kk] for kk in ("idx_sha256", "dist_sha256", "recall_at_k")}), flush=True)
if "json" in opts:
    json.dump(out, open(opts["json"], "w"), indent=1)
