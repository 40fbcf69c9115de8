"""A/B harness for the schedule switches of libnnd_b200 (none may change a result).

    python scripts/r2_ab.py [N] [d] [k] [metric] [--configs name,name,...] [--reps R]

Builds the same graph once per configuration in a FRESH subprocess (the switches are read by nnd_create / the engine from
the environment), with a hard per-configuration timeout so that a hung experimental kernel cannot take the GPU box down
with it, and prints for each: bit-equality of (indices, distances, per-iteration accepted pushes) against the default
schedule, build ms, the local-join kernel ms (CUDA events inside the library), host stage laps and the in-library stage timers.
Round 2's outcomes: DESIGN.md section 6 (ledger), profiles/ab_r02_*.json.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONFIGS = {
    "default": {},
    "noapplyshort": {"NND_APPLY_SHORT_SORT": "0"},
    "nocandskip": {"NND_CAND_SKIP": "0"},
    "candpartition": {"NND_CAND_PARTITION": "1"},
    "forestpm96": {"NND_FOREST_PM_MB": "96"},
    "noforestpm": {"NND_FOREST_PM": "0"},
    "norank": {"NND_RANK_SCATTER": "0"},
    "noleafsplit": {"NND_LEAF_SPLIT": "0"},
    "nocarveout": {"NND_SMEM_CARVEOUT": "0"},
    "round1-defaults": {"NND_RANK_SCATTER": "0", "NND_LEAF_SPLIT": "0", "NND_SMEM_CARVEOUT": "0", "NND_FOREST_PM": "0", "NND_CAND_SKIP": "0",
                        "NND_APPLY_SHORT_SORT": "0"},
}


def child(n, d, k, metric, reps, out_path):
    import hashlib
    import time
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    from bench import latent
    from pynndescent_b200 import distances as pdist
    from pynndescent_b200.engine import Engine
    cache = "/dev/shm/nnd_ab_%d_%d.npy" % (n, d)          # the generator needs half a minute at 10M points: share it between the children
    if os.path.exists(cache):
        X = np.load(cache)
    else:
        X = latent(n, d, seed=0)
        try:
            np.save(cache, X)
        except OSError:
            pass
    metric_id, dev_corr, host_corr, angular = pdist.resolve(metric)
    n_trees = max(3, min(12, int(round(2.0 * np.log10(n))))); n_iters = max(5, int(round(np.log2(n))))
    leaf_size = max(60, min(256, 5 * k)); C_ = min(60, k)
    eng = Engine(X, metric_id, k)
    eng.stage_profile(True)
    res = {}
    for rep in range(reps + 1):            # first build = warm-up
        rs = np.random.RandomState(42)
        rng = rs.randint(-2 ** 31 + 1, 2 ** 31 - 1, 3).astype(np.int64); rs.randint(-2 ** 31 + 1, 2 ** 31 - 1, 3)
        trees = rs.randint(-2 ** 31 + 1, 2 ** 31 - 1, size=(n_trees, 3)).astype(np.int64)
        eng.profile(reset=True); eng.stage_ms(reset=True)
        stage = {}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        idx_t, dist_t, iters, c_hist, stats = eng.build(rng, trees, leaf_size, angular, C_, 8, n_iters, 0.001, dev_corr, timings=stage)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        prof = eng.profile(reset=True); dev = eng.stage_ms(reset=True)
        gi, gd = idx_t.cpu().numpy(), dist_t.cpu().numpy()
        res = {"build_ms": ms, "join_ms": prof["join"]["ms"], "join_launches": prof["join"]["launches"],
               "leaf_ms": prof["leaf"]["ms"], "iters": iters, "c_hist": list(map(int, c_hist)),
               "stage_ms": {kk: round(vv * 1e3, 2) for kk, vv in stage.items()}, "dev_ms": {kk: round(vv, 2) for kk, vv in dev.items() if vv > 0.005},
               "idx_sha": hashlib.sha256(gi.tobytes()).hexdigest(), "dist_sha": hashlib.sha256(gd.tobytes()).hexdigest()}
    json.dump(res, open(out_path, "w"))


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    opts = dict(a[2:].split("=", 1) for a in sys.argv[1:] if a.startswith("--") and "=" in a)
    if args and args[0] == "child":
        return child(int(args[1]), int(args[2]), int(args[3]), args[4], int(args[5]), args[6])
    n = int(args[0]) if len(args) > 0 else 1_000_000
    d = int(args[1]) if len(args) > 1 else 128
    k = int(args[2]) if len(args) > 2 else 15
    metric = args[3] if len(args) > 3 else "euclidean"
    reps = int(opts.get("reps", 2))
    names = opts["configs"].split(",") if "configs" in opts else list(CONFIGS)
    tmo = int(opts.get("timeout", 240))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    base = None
    rows = []
    for name in names:
        env = {kk: vv for kk, vv in os.environ.items() if not kk.startswith("NND_")}
        env.update(CONFIGS[name])
        out = os.path.join(ROOT, "gpurun_out", f"ab_{name}.json")
        if os.path.exists(out):
            os.remove(out)
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "child", str(n), str(d), str(k), metric, str(reps), out],
                               env=env, timeout=tmo, capture_output=True, text=True)
            status = "ok" if p.returncode == 0 and os.path.exists(out) else f"FAILED rc={p.returncode}: {p.stderr.strip()[-400:]}"
        except subprocess.TimeoutExpired:
            status = f"TIMEOUT after {tmo} s (hung kernel?)"
        if not status.startswith("ok"):
            print(f"{name:16s} {status}", flush=True)
            rows.append({"config": name, "status": status})
            continue
        r = json.load(open(out))
        if name == "default" or base is None:
            base = base or r
        same = (r["idx_sha"] == base["idx_sha"], r["dist_sha"] == base["dist_sha"], r["c_hist"] == base["c_hist"])
        print(f"{name:16s} build {r['build_ms']:8.1f} ms  join {r['join_ms']:7.2f} ms ({int(r['join_launches'])} launches)  leaf {r['leaf_ms']:6.2f} ms  "
              f"bit-equal idx/dist/c: {same[0]}/{same[1]}/{same[2]}  stages {r['stage_ms']}\n{'':17s}device events {r.get('dev_ms')}", flush=True)
        r.update({"config": name, "status": "ok", "equal_to_default": all(same)})
        rows.append(r)
    # ---- compose the winners: every switch set that is bit-equal and at least 1 % faster than the default on its own goes
    # into one combined configuration, which is then run like the others
    by = {r["config"]: r for r in rows if r.get("status") == "ok"}
    if "default" in by and "auto" not in opts.get("skip", ""):
        t0 = by["default"]["build_ms"]
        combined = {}
        for name, r in by.items():
            if name in ("default", "all") or not r.get("equal_to_default") or r["build_ms"] > 0.99 * t0:
                continue
            if any(kk in combined and combined[kk] != vv for kk, vv in CONFIGS[name].items()):
                continue                                     # conflicting values of a tuning knob: keep the first (faster rows come first below)
            combined.update(CONFIGS[name])
        if combined:
            CONFIGS["auto-combined"] = combined
            env = {kk: vv for kk, vv in os.environ.items() if not kk.startswith("NND_")}; env.update(combined)
            out = os.path.join(ROOT, "gpurun_out", "ab_auto-combined.json")
            try:
                p = subprocess.run([sys.executable, os.path.abspath(__file__), "child", str(n), str(d), str(k), metric, str(reps), out],
                                   env=env, timeout=tmo, capture_output=True, text=True)
                if p.returncode == 0 and os.path.exists(out):
                    r = json.load(open(out))
                    same = r["idx_sha"] == by["default"]["idx_sha"] and r["dist_sha"] == by["default"]["dist_sha"] and r["c_hist"] == by["default"]["c_hist"]
                    print(f"auto-combined {combined}\n                 build {r['build_ms']:8.1f} ms (default {t0:.1f})  join {r['join_ms']:7.2f} ms  leaf {r['leaf_ms']:6.2f} ms  "
                          f"bit-equal: {same}  stages {r['stage_ms']}", flush=True)
                    r.update({"config": "auto-combined", "env": combined, "status": "ok", "equal_to_default": same}); rows.append(r)
                else:
                    print("auto-combined FAILED:", p.stderr.strip()[-400:], flush=True)
            except subprocess.TimeoutExpired:
                print("auto-combined TIMEOUT", flush=True)
    json.dump({"n": n, "d": d, "k": k, "metric": metric, "rows": rows}, open(os.path.join(ROOT, "gpurun_out", "ab_summary.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
