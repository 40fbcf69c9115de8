"""Golden fixtures for `inner_product` and `true_angular` (fast_distance_alternatives, distances.py:2176-2184), generated from
the REAL reference (build container only):

    python tests/golden/make_golden_metrics2.py

* kat_metrics2.npz      -- pairwise matrices of the reference's alternative_inner_product on mixed-sign rows (incl. a zero row,
                           duplicates, orthogonal pairs -> FLOAT32_MAX sentinel), its correction, and true_angular's correction of
                           the reference's alternative_cosine matrix
* latent_inner_1500x20  -- staged build replay (make_golden.staged) for metric="inner_product" on mixed-sign data: about half of
                           all pairs have a non-positive inner product and carry the sentinel distance
* latent_trueang_1500x20 -- the same for metric="true_angular" (alternative_cosine kernel, euclidean trees, no normalisation)
* latent_minkowski3_1500x20 -- the same for metric="minkowski", metric_kwds={"p": 3.0}
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, latent, staged, pd  # noqa: E402  (imports the reference through the shim, T=8)


def kat():
    out = {}
    rs = np.random.RandomState(41)
    A = rs.standard_normal((24, 37)).astype(np.float32)
    A[5] = 0.0; A[6] = A[7]; A[8] = 2.5; A[9, ::2] = 0.0; A[10] = -A[11]
    A[12] = 0.0; A[12, 0] = 1.0; A[13] = 0.0; A[13, 1] = 1.0            # orthogonal pair: inner product exactly 0
    out["real_A"] = A
    out["dist_inner_product"] = np.array([[pd.alternative_inner_product(A[i], A[j]) for j in range(24)] for i in range(24)], np.float32)
    out["corr_inner_product"] = pd.correct_alternative_inner_product(out["dist_inner_product"])
    out["true_inner_product"] = np.array([[pd.inner_product(A[i], A[j]) for j in range(24)] for i in range(24)], np.float32)
    out["dist_true_angular"] = np.array([[pd.alternative_cosine(A[i], A[j]) for j in range(24)] for i in range(24)], np.float32)
    out["corr_true_angular"] = pd.true_angular_from_alt_cosine(out["dist_true_angular"])
    out["true_true_angular"] = np.array([[pd.true_angular(A[i], A[j]) for j in range(24)] for i in range(24)], np.float32)
    # minkowski(x, y, p) (distances.py:138-153; p reaches the function positionally from metric_kwds.values())
    for p in (0.5, 1.5, 3.0, 4.0):
        out["dist_minkowski_p%g" % p] = np.array([[pd.minkowski(A[i], A[j], p) for j in range(24)] for i in range(24)], np.float32)
    np.savez_compressed(os.path.join(OUT, "kat_metrics2.npz"), **out)


if __name__ == "__main__":
    kat()
    staged("latent_inner_1500x20", latent(1500, 20, 31), "inner_product", 10, 13)
    staged("latent_trueang_1500x20", latent(1500, 20, 32), "true_angular", 10, 14)
    staged("latent_minkowski3_1500x20", latent(1500, 20, 33), "minkowski", 10, 15, metric_kwds={"p": 3.0})
