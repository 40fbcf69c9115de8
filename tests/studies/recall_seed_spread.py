"""Seed-to-seed spread of recall@15 on the cfg2 workload (1M x 128, latent-16) for the two forest rng schemes of the oracle:
"dfs" (= the reference's forest, rp_trees.py:2177-2302) and "level" (= the GPU forest's).  Same data, same 1000-row sample
(the exact neighbours are read from tests/golden/cfg2_full_oracle.json), random_state = 42, 1, 2, 3, 4.

    python tests/studies/recall_seed_spread.py            -> tests/studies/recall_seed_spread_cfg2.json

The north_star bar "recall within 0.5 pt of the reference on the same seed" compares two DIFFERENT random forests (the reference's
rng stream is tied to its depth-first recursion); this study shows how much recall moves when only the seed changes."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import latent
from oracle import nnd_oracle as orc

fx = json.load(open(os.path.join(ROOT, "tests", "golden", "cfg2_full_oracle.json")))
n, d, k = fx["n"], fx["d"], fx["k"]
true = np.asarray(fx["exact_knn_of_sample"])
rows = np.random.RandomState(0).choice(n, 1000, replace=False)
orc.lib().orc_set_num_threads(os.cpu_count())
X = latent(n, d, 0)
out = {"workload": "cfg2 1Mx128 euclidean k=15", "rows": "the fixture's 1000-row sample", "runs": []}
for seed in (42, 1, 2, 3, 4):
    for scheme in ("dfs", "level"):
        t0 = time.time()
        o = orc.OracleNNDescent(X, "euclidean", n_neighbors=k, random_state=seed, n_threads=8, forest=scheme)
        gi, _ = o.neighbor_graph
        rec = float(np.mean([len(np.intersect1d(true[i], gi[r])) for i, r in enumerate(rows)]) / k)
        out["runs"].append({"seed": seed, "forest": scheme, "recall_at_15": rec, "iterations": int(o.n_iters_run), "seconds": round(time.time() - t0, 1)})
        print(out["runs"][-1], flush=True)
for scheme in ("dfs", "level"):
    r = np.array([x["recall_at_15"] for x in out["runs"] if x["forest"] == scheme])
    out[scheme] = {"mean": float(r.mean()), "std": float(r.std(ddof=1)), "min": float(r.min()), "max": float(r.max())}
json.dump(out, open(os.path.join(ROOT, "tests", "studies", "recall_seed_spread_cfg2.json"), "w"), indent=1)
print(out["dfs"], out["level"])
