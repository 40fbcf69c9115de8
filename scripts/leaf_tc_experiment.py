"""Leaf join on tensor cores -- the measurement that settles north_star's "euclidean / cosine / dot on tensor cores" clause for the
one stage where it could matter (leaf join: arithmetic intensity 10-37 flop/B; the local join is a gather).

    python scripts/leaf_tc_experiment.py [N] [--exact]      (default N = 1,000,000; d = 128, k = 15, 12 trees, leaf_size 75)

Tree by tree, BEFORE the product's leaf init of that tree runs, the experiment kernel (csrc/x_leaf_tc.cu: BF16 mma.sync Gram of
every leaf -> conservative lower bound of d^2 -> compare with the current heap roots) counts how many leaf pairs a tensor-core
prefilter could NOT exclude; those would still need the exact FP32 difference-form chain, because heaps must hold the canonical
distances bit for bit.  Printed per tree: pairs, survivor fraction, (with --exact) the fraction that truly passes, the Gram
kernel's time, and the product's exact leaf-distance kernel + push kernel times for the same tree.
Writes gpurun_out/leaf_tc_experiment.json.
"""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import latent
from pynndescent_b200 import distances as pdist
from pynndescent_b200._ffi import check, lib
from pynndescent_b200.engine import Engine

args = [a for a in sys.argv[1:] if not a.startswith("--")]
N = int(args[0]) if args else 1_000_000
exact = "--exact" in sys.argv
d, k, leaf_size, n_trees = 128, 15, 75, max(3, min(12, int(round(2.0 * np.log10(N)))))
X = latent(N, d, 0)
eng = Engine(X, pdist.resolve("euclidean")[0], k)
rs = np.random.RandomState(42)
rs.randint(-2 ** 31 + 1, 2 ** 31 - 1, 3); rs.randint(-2 ** 31 + 1, 2 ** 31 - 1, 3)
trees = rs.randint(-2 ** 31 + 1, 2 ** 31 - 1, size=(n_trees, 3)).astype(np.int64)
lp, lo, n_leaves, max_leaf = eng.forest(trees, leaf_size, False)
lo_h = lo.cpu().numpy()
cuts = [int(np.searchsorted(lo_h, t * N)) for t in range(n_trees + 1)]
cap = eng.default_push_capacity()
eng.workspace(1, 1, cap)
eng.make_heap()
eng.stage_profile(True)
counts = torch.zeros(3, dtype=torch.int64, device=eng.device)
rows = []
for t in range(n_trees):
    l0, l1 = cuts[t], cuts[t + 1]
    for rep in range(2):                                         # second launch = warm timing
        counts.zero_()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(eng.stream)
        check(lib.nnd_x_leaf_gram_tc(eng._ctx, C.byref(eng.problem()), C.c_void_p(lp.data_ptr()), C.c_void_p(lo.data_ptr()), l0, l1, max_leaf,
                                     int(exact and rep == 1), C.c_void_p(counts.data_ptr()), C.c_void_p(eng.stream.cuda_stream)))
        ev1.record(eng.stream); torch.cuda.synchronize()
        if rep == 0:
            ms_tc = ev0.elapsed_time(ev1)
    if not exact:
        ms_tc = ev0.elapsed_time(ev1)
    pairs, surv, passed = [int(v) for v in counts.cpu()]
    eng.stage_ms(reset=True)
    eng.leaf_fused_range(lp, lo, l0, l1, max_leaf, pos0=int(lo_h[l0]), npos=int(lo_h[l1] - lo_h[l0]), cap=cap)   # the product's leaf init of this tree
    torch.cuda.synchronize()
    sm = eng.stage_ms(reset=True)
    row = {"tree": t, "leaves": l1 - l0, "pairs": pairs, "bf16_survivors": surv, "survivor_fraction": surv / max(pairs, 1),
           "exact_pass_fraction": (passed / max(pairs, 1)) if exact else None, "conservative": (passed <= surv) if exact else None,
           "tc_gram_ms": round(ms_tc, 3), "product_leaf_dist_ms": round(sm["leaf_dist"], 3), "product_leaf_push_ms": round(sm["leaf_push"], 3)}
    rows.append(row); print(row, flush=True)
tot_pairs = sum(r["pairs"] for r in rows); tot_surv = sum(r["bf16_survivors"] for r in rows)
summary = {"n": N, "d": d, "k": k, "trees": n_trees, "rows": rows, "pairs": tot_pairs, "survivor_fraction_all_trees": tot_surv / tot_pairs,
           "tc_gram_ms_total": round(sum(r["tc_gram_ms"] for r in rows), 3), "product_leaf_dist_ms_total": round(sum(r["product_leaf_dist_ms"] for r in rows), 3),
           "product_leaf_push_ms_total": round(sum(r["product_leaf_push_ms"] for r in rows), 3)}
print({kk: vv for kk, vv in summary.items() if kk != "rows"}, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(summary, open("gpurun_out/leaf_tc_experiment.json", "w"), indent=1)
