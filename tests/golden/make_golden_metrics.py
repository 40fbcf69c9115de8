"""Golden fixtures for the SURVEY 8f-2 metrics, generated from the REAL reference (build container only):

    python tests/golden/make_golden_metrics.py

* kat_metrics.npz   -- pairwise matrices of the reference's own numba distance functions on small inputs of the
                       right domain (real / categorical / binary / non-negative), incl. zero rows and duplicates
* <case>.npz        -- the staged build replay of make_golden.staged() (leaf_array, heaps after each stage,
                       candidates, c per iteration, final graph) for one case per accumulation class
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, latent, staged, pd  # noqa: E402  (imports the reference through the shim, T=8)

REAL = ("manhattan", "chebyshev", "canberra")
BINARY = {"jaccard": pd.alternative_jaccard, "dice": pd.dice, "matching": pd.matching, "kulsinski": pd.kulsinski,
          "rogerstanimoto": pd.rogers_tanimoto, "sokalmichener": pd.sokal_michener, "russellrao": pd.russellrao,
          "sokalsneath": pd.sokal_sneath, "yule": pd.yule}


def binary_data(n, d, seed):
    return (latent(n, d, seed) > 0.35).astype(np.float32)


def categorical_data(n, d, seed):
    return np.clip(np.round(latent(n, d, seed) * 1.2), -2, 2).astype(np.float32)


def nonneg_data(n, d, seed):
    X = np.abs(latent(n, d, seed)).astype(np.float32)
    X[X < 0.15] = 0.0
    return X


def kat():
    out = {}
    rs = np.random.RandomState(11)
    A = rs.standard_normal((24, 37)).astype(np.float32)
    A[5] = 0.0; A[6] = A[7]; A[8] = 2.5; A[9, ::2] = 0.0
    out["real_A"] = A
    for name, fn in (("manhattan", pd.manhattan), ("chebyshev", pd.chebyshev), ("canberra", pd.canberra)):
        out[f"dist_{name}"] = np.array([[fn(A[i], A[j]) for j in range(24)] for i in range(24)], np.float32)
    Cat = categorical_data(24, 37, 12); Cat[5] = 0.0; Cat[6] = Cat[7]
    out["cat_A"] = Cat
    out["dist_hamming"] = np.array([[pd.hamming(Cat[i], Cat[j]) for j in range(24)] for i in range(24)], np.float32)
    B = binary_data(24, 37, 13); B[5] = 0.0; B[6] = B[7]; B[8] = 1.0
    out["bin_A"] = B
    for name, fn in BINARY.items():
        out[f"dist_{name}"] = np.array([[fn(B[i], B[j]) for j in range(24)] for i in range(24)], np.float32)
    out["corr_jaccard"] = pd.correct_alternative_jaccard(out["dist_jaccard"])
    H = nonneg_data(24, 37, 14); H[5] = 0.0; H[6] = H[7]
    out["nonneg_A"] = H
    out["dist_hellinger"] = np.array([[pd.alternative_hellinger(H[i], H[j]) for j in range(24)] for i in range(24)], np.float32)
    out["corr_hellinger"] = pd.correct_alternative_hellinger(out["dist_hellinger"])
    np.savez_compressed(os.path.join(OUT, "kat_metrics.npz"), **out)


if __name__ == "__main__":
    kat()
    staged("latent_manhattan_1500x20", latent(1500, 20, 21), "manhattan", 10, 3)
    staged("latent_chebyshev_1200x12", latent(1200, 12, 22), "chebyshev", 10, 4)
    staged("binary_jaccard_1200x40", binary_data(1200, 40, 23), "jaccard", 10, 5)
    staged("cat_hamming_1200x30", categorical_data(1200, 30, 24), "hamming", 10, 6)
    # alternative_hellinger accumulates in float32 under fastmath (distances.py:1370-1386 `locals`): its low bits
    # depend on LLVM's summation order, and a rounding-level tie between two distances can swap two slots of a heap
    # row (same final graph).  Of five (data seed, index seed) pairs tried, (25, 9) and (26, 9) each showed ONE such
    # swap in one row against the oracle's sequential order; (27, 10), (28, 11), (29, 12) are identical in every
    # stage -- the first of those is the committed fixture.
    staged("nonneg_hellinger_1200x24", nonneg_data(1200, 24, 27), "hellinger", 10, 10)
