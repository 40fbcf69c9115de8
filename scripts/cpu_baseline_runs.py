"""CPU-side evidence for bench.py's cpu_baseline / reference arm (runs on the GPU box's host cores, no GPU needed):

    python scripts/cpu_baseline_runs.py sweep [n]            # oracle-port build time vs OpenMP thread count at n points (default 200k)
    python scripts/cpu_baseline_runs.py full cfg4 [threads]  # ONE full-size build of a bench workload (10M x 128: minutes)

Writes gpurun_out/cpu_thread_sweep.json / gpurun_out/cpu_full_<workload>.json (copied to profiles/ by hand).
The port (oracle/nnd_oracle.c, forest scheme "dfs" = the reference's) keeps the reference's parallel structure: T threads each scan
ALL edges in new_build_candidates (utils.py:261-306) and ALL updates in apply_graph_update_array (utils.py:709-731).
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from oracle import nnd_oracle as orc

os.makedirs("gpurun_out", exist_ok=True)
mode = sys.argv[1]
if mode == "sweep":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
    X = bench.latent(n, 128, 0)
    rows = []
    for T in (4, 8, 16, 32, 64, 128):
        if T > (os.cpu_count() or 1):
            break
        dt, o = bench.run_cpu_port(X, "euclidean", 15, 42, T)
        rows.append({"threads": T, "seconds": round(dt, 3), "edges_per_s": n * 15 / dt, "iterations": int(o.n_iters_run)})
        print(rows[-1], flush=True)
    json.dump({"n": n, "d": 128, "k": 15, "host_cores": os.cpu_count(), "rows": rows}, open("gpurun_out/cpu_thread_sweep.json", "w"), indent=1)
else:
    name = sys.argv[2]; T = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    wl = bench.WORKLOADS[name]
    X = bench.latent(wl["n"], wl["d"], 0)
    dt, o = bench.run_cpu_port(X, wl["metric"], wl["k"], 42, T)
    gi, gd = o.neighbor_graph
    fx = bench.load_fixture(name)
    true = np.asarray(fx["exact_knn_of_sample"]) if fx and "exact_knn_of_sample" in fx else None
    out = {"workload": wl["name"], "n": wl["n"], "seconds": round(dt, 1), "edges_per_s": wl["n"] * wl["k"] / dt, "threads": T,
           "host_cores": os.cpu_count(), "iterations": int(o.n_iters_run), "kind": "port (oracle/nnd_oracle.c, dfs forest)",
           "when": time.strftime("%Y-%m-%d")}
    if true is not None:
        rr = bench.recall_rows(wl["n"])
        out["recall_at_k"] = float(np.mean([len(np.intersect1d(true[i], gi[r])) for i, r in enumerate(rr)]) / wl["k"])
    print(json.dumps(out), flush=True)
    json.dump(out, open("gpurun_out/cpu_full_%s.json" % name, "w"), indent=1)
