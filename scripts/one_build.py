"""One warm + one measured build (ncu target): python scripts/one_build.py N d k [metric] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import latent
from pynndescent_b200 import NNDescent
N, d, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
metric = sys.argv[4] if len(sys.argv) > 4 else "euclidean"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 1
X = latent(N, d)
for _ in range(reps):
    idx = NNDescent(X, metric, n_neighbors=k, random_state=42)
torch.cuda.synchronize()
print("done", idx.timings)
