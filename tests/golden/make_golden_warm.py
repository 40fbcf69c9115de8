"""Golden fixtures for the init_graph / init_dist warm start (pynndescent_.py:1225-1242, utils.py:836-860), from
the REAL reference (build container only):   python tests/golden/make_golden_warm.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, latent, pynndescent  # noqa: E402  (reference through the shim, numba threads T=8)


def main():
    X = latent(1500, 16, 31)
    k = 10
    rs = np.random.RandomState(4)
    # a deliberately poor start: random neighbours, ragged (-1 holes), a duplicate and a self loop per row
    g_rand = rs.randint(0, 1500, size=(1500, 7)).astype(np.int32)
    g_rand[:, 3] = g_rand[:, 1]
    g_rand[::5, 5] = -1
    g_rand[:, 0] = np.arange(1500)
    # a good start: the exact 12-NN graph (wider than k) with TRUE euclidean distances as init_dist -- the reference
    # pushes them verbatim into a heap whose other entries are squared distances (its own quirk, reproduced)
    D = ((X[:, None, :].astype(np.float64) - X[None, :, :]) ** 2).sum(-1)
    g_knn = np.argsort(D, 1, kind="stable")[:, :12].astype(np.int32)
    d_knn = np.sqrt(np.take_along_axis(D, g_knn, 1)).astype(np.float32)
    out = {"X": X, "k": k, "seed": 21, "T": 8, "g_rand": g_rand, "g_knn": g_knn, "d_knn": d_knn}
    for name, kw in (("rand", dict(init_graph=g_rand)), ("knn", dict(init_graph=g_knn)),
                     ("knn_dist", dict(init_graph=g_knn, init_dist=d_knn))):
        for metric in ("euclidean", "cosine"):
            idx = pynndescent.NNDescent(X, metric, n_neighbors=k, random_state=21, **kw)
            gi, gd = idx.neighbor_graph
            out[f"{name}_{metric}_idx"], out[f"{name}_{metric}_dist"] = gi, gd
            print(name, metric, "self at col 0:", float((gi[:, 0] == np.arange(1500)).mean()))
    np.savez_compressed(os.path.join(OUT, "warm_start_1500x16.npz"), **out)


if __name__ == "__main__":
    main()
