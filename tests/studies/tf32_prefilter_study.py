"""Design study (CPU, oracle-driven; not a test): how much of the local join's EXACT fp32 work would a conservative
tensor-core prefilter remove?

The local join must decide `d < thr` on bit-exact distances, so a TF32 / BF16 Gram matrix can
never REPLACE the diff-form chain -- but it can bound it: with inputs rounded to t mantissa bits,
|G_t - G| <= (2^-(t) + eps) * |x||y|, hence  d2_approx - 2^-(t-1)*1.02*|x||y|  <=  d2_exact.  A pair whose lower bound
already fails both thresholds cannot be emitted and needs no exact evaluation.  This script replays an NN-descent build
with the oracle, and for a row sample of every iteration reports: pairs evaluated, fraction emitted (exact), fraction that
survives the TF32 (t = 11) and BF16 (t = 8) lower-bound tests.  Survivors / pairs = the share of today's fp32 work that
would remain.

    python tests/studies/tf32_prefilter_study.py [N] [d] [k] [rows_sampled]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import latent  # noqa: E402
from oracle import nnd_oracle as orc  # noqa: E402


def round_mantissa(x, bits):
    """round-to-nearest-even to `bits` explicit mantissa bits (TF32: 10, BF16: 7)."""
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    drop = 23 - bits
    u = u + ((1 << (drop - 1)) - 1) + ((u >> drop) & 1)
    u &= ~np.uint64((1 << drop) - 1)
    return u.astype(np.uint32).view(np.float32)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 15
    ns = int(sys.argv[4]) if len(sys.argv) > 4 else 4000
    X = latent(n, d, 0)
    n2 = (X.astype(np.float64) ** 2).sum(1)
    Xt = {"tf32": round_mantissa(X, 10).astype(np.float64), "bf16": round_mantissa(X, 7).astype(np.float64)}
    unit = {"tf32": 2.0 ** -11, "bf16": 2.0 ** -8}
    from sklearn.utils import check_random_state
    rs = check_random_state(42)
    rng_state = rs.randint(orc.INT32_MIN, orc.INT32_MAX, 3).astype(np.int64); rs.randint(orc.INT32_MIN, orc.INT32_MAX, 3)
    n_trees = max(3, min(12, int(round(2.0 * np.log10(n))))); n_iters = max(5, int(round(np.log2(n))))
    leaf_array, _ = orc.make_forest(X, k, n_trees, None, rs, False, 200, "level")
    heap = orc.make_heap(n, k)
    orc.init_rp_tree(X, "euclidean", heap, leaf_array)
    orc.init_random(X, "euclidean", heap, rng_state)
    C = min(60, k)
    rows = np.sort(np.random.RandomState(0).choice(n, ns, replace=False))
    print(f"{n}x{d} k={k}: per iteration over {ns} sampled rows")
    print(" it      pairs  emitted  survive_tf32  survive_bf16   c/(kN)")
    tot = np.zeros(4)
    for it in range(n_iters):
        newc, oldc = orc.new_build_candidates(heap, C, rng_state, 8)
        thr = heap[1][:, 0].astype(np.float64)
        pa, pb = [], []
        for r in rows:
            nw = newc[r][newc[r] >= 0]; ol = oldc[r][oldc[r] >= 0]
            if len(nw) == 0:
                continue
            iu = np.triu_indices(len(nw))
            pa.append(nw[iu[0]]); pb.append(nw[iu[1]])
            if len(ol):
                pa.append(np.repeat(nw, len(ol))); pb.append(np.tile(ol, len(nw)))
        pa = np.concatenate(pa); pb = np.concatenate(pb)
        exact = ((X[pa].astype(np.float64) - X[pb].astype(np.float64)) ** 2).sum(1)
        lim = np.maximum(thr[pa], thr[pb])
        emitted = exact < lim
        row = [len(pa), emitted.mean()]
        for name in ("tf32", "bf16"):
            G = (Xt[name][pa] * Xt[name][pb]).sum(1)
            approx = n2[pa] + n2[pb] - 2.0 * G
            bound = 2.0 * (2.0 * unit[name] * 1.02) * np.sqrt(n2[pa] * n2[pb])
            survive = (approx - bound) < lim
            assert not np.any(emitted & ~survive), "the lower bound dropped a pair that must be emitted"
            row.append(survive.mean())
        c, _ = orc.process_candidates(X, "euclidean", heap, newc, oldc, 8)
        print(f" {it:2d} {row[0]:10d}  {row[1]:7.3f}  {row[2]:12.3f}  {row[3]:12.3f}  {c / (k * n):7.4f}")
        tot += np.array([row[0], row[0] * row[1], row[0] * row[2], row[0] * row[3]])
        if c <= 0.001 * k * n:
            break
    print(f"all iterations: emitted {tot[1] / tot[0]:.3f} of pairs; exact work left after TF32 filter {tot[2] / tot[0]:.3f}, "
          f"after BF16 filter {tot[3] / tot[0]:.3f}")


if __name__ == "__main__":
    main()
