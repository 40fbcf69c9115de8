"""Stage timings of one build (dev tool): python scripts/time_build.py N d k [metric]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pynndescent_b200 import NNDescent

def latent(n, d, seed=0, latent_dim=16):
    rs = np.random.RandomState(seed)
    Z = rs.standard_normal((n, latent_dim)).astype(np.float32)
    A = rs.standard_normal((latent_dim, d)).astype(np.float32) / np.sqrt(latent_dim)
    return (Z @ A + 0.05 * rs.standard_normal((n, d)).astype(np.float32)).astype(np.float32)

N, d, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
metric = sys.argv[4] if len(sys.argv) > 4 else "euclidean"
X = latent(N, d)
for rep in range(2):
    t0 = time.perf_counter()
    idx = NNDescent(X, metric, n_neighbors=k, random_state=42)
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    print(f"rep {rep}: total {t:.3f}s  edges/s {N*k/t/1e6:.2f}M  timings {idx.timings} iters {idx.n_iters_run} c {idx.c_hist}")
    print("   join stats [rows gathered, pairs, pushes, blocks]:", idx.join_stats, "launches", idx._engine.launch_count())
gi, gd = idx.neighbor_graph
rows = np.random.RandomState(0).choice(N, 1000, replace=False)
Xr = X[rows].astype(np.float64)
d2 = (Xr**2).sum(1)[:, None] + (X.astype(np.float64)**2).sum(1)[None, :] - 2 * Xr @ X.astype(np.float64).T
true = np.argsort(d2, 1)[:, :k]
print("recall", np.mean([len(np.intersect1d(true[i], gi[r])) / k for i, r in enumerate(rows)]))
