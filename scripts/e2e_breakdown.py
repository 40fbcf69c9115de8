import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import latent
from pynndescent_b200 import NNDescent
X = latent(1000000, 128)
Xp = torch.from_numpy(X).pin_memory().numpy()
print("pinned view is_pinned:", torch.from_numpy(Xp).is_pinned())
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    idx = NNDescent(Xp, "euclidean", n_neighbors=15, random_state=42)
    t1 = time.perf_counter()
    g = idx.neighbor_graph
    t2 = time.perf_counter()
    print(f"rep{rep}: ctor {t1-t0:.3f}s neighbor_graph {t2-t1:.3f}s timings", {k: round(v*1e3,1) for k,v in idx.timings.items()})
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
idx = NNDescent(Xp, "euclidean", n_neighbors=15, random_state=42); g = idx.neighbor_graph
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
