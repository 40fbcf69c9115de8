"""Query path timing (secondary metric): build -> prepare -> query on one GPU.
    python scripts/query_bench.py N d k nq [metric]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import latent
from pynndescent_b200 import NNDescent
N, d, k, nq = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
metric = sys.argv[5] if len(sys.argv) > 5 else "euclidean"
both = latent(N + nq, d, 0)
X, Q = both[:N], both[N:]
t0 = time.perf_counter(); idx = NNDescent(X, metric, n_neighbors=k, random_state=42); torch.cuda.synchronize(); t1 = time.perf_counter()
idx.prepare(); torch.cuda.synchronize(); t2 = time.perf_counter()
qi, qd = idx.query(Q[:1000], k=10, epsilon=0.1)   # warm
torch.cuda.synchronize(); t3 = time.perf_counter()
qi, qd = idx.query(Q, k=10, epsilon=0.1)
torch.cuda.synchronize(); t4 = time.perf_counter()
st = idx._search_gpu.last_stats
rows = np.random.RandomState(0).choice(nq, min(nq, 500), replace=False)
Xd = torch.from_numpy(X).cuda().double(); Qd = torch.from_numpy(Q[rows]).cuda().double()
if metric == "correlation":      # correlation distance = cosine distance of the row-centred vectors
    Xd = Xd - Xd.mean(dim=1, keepdim=True); Qd = Qd - Qd.mean(dim=1, keepdim=True)
dm = (Qd * Qd).sum(1)[:, None] + (Xd * Xd).sum(1)[None, :] - 2 * Qd @ Xd.T if metric == "euclidean" else \
    1 - (Qd / Qd.norm(dim=1, keepdim=True)) @ (Xd / Xd.norm(dim=1, keepdim=True)).T
true = dm.topk(10, dim=1, largest=False).indices.cpu().numpy()
rec = np.mean([len(np.intersect1d(true[i], qi[r])) / 10 for i, r in enumerate(rows)])
print(f"N={N} d={d} k={k} {metric}: build {t1-t0:.3f}s prepare {t2-t1:.3f}s query {nq} pts {t4-t3:.3f}s = {nq/(t4-t3):.0f} QPS "
      f"recall@10 {rec:.3f} stats {st} ({st['distance_evals']/nq:.0f} dist evals/query)")
