#!/usr/bin/env python
"""bench.py -- kNN-graph build edges/sec (BASELINE.json metric) on synthetic float32 data.

    python bench.py --gpus N --steps K --warmup W              # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...     # CPU port of the reference path (rank 0 only)
    python bench.py --workload cfg2|cfg3|cfg5 ...               # the other BASELINE.json configurations
    python bench.py --impl reference --thread-sweep | --full-size  # CPU-side evidence for the baseline (thread count; one whole build)

Default workload = the configuration BASELINE.json quotes the metric on: ONE graph over 10M x 128 float32 points,
k = 15, euclidean ("cfg4").  It fits one B200 (X 5.1 GB + heaps 1.35 GB + scratch), so `--gpus 1` builds it on one GPU and
`--gpus N` shards THE SAME graph over N GPUs (strong scaling; N = 8 is BASELINE.json configs[3]).

A "step" is one whole build (RP forest -> leaf init -> init_random -> NN-descent -> deheap_sort) with the data matrix
already resident in HBM (`value`), and the same build through the public NNDescent(X_host) API with the H2D copy of X and
the D2H copy of the graph inside the timed region (`e2e`).  Device time is measured with CUDA events on the launching
stream; N>1 = one process per GPU, max over ranks.  One JSON line is printed by rank 0.

`parity`: sha256 of the graph of the last timed build (indices, distances) and the per-iteration accepted-push counts --
identical for every N by construction, and equal to the CPU oracle's graph when a full-size fixture exists
(tests/golden/<workload>_full_oracle.json, written by tests/golden/make_full_size_checksum.py); `recall_at_k` is measured on
the fixture's fixed 1000-row sample next to the dfs-forest (= reference scheme) oracle's recall on the same rows.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[3] / the metric's own configuration ("edges/sec at N=10M d=128 k=15")
    "cfg4": dict(n=10_000_000, d=128, k=15, metric="euclidean", name="10Mx128 float32 euclidean k=15 (latent-16 synthetic)"),
    "cfg2": dict(n=1_000_000, d=128, k=15, metric="euclidean", name="1Mx128 float32 euclidean k=15 (SIFT-shape latent-16 synthetic)"),
    "cfg3": dict(n=2_000_000, d=256, k=15, metric="cosine", name="2Mx256 float32 cosine k=15 (NYTimes-shape latent-16 synthetic)"),
    "cfg5": dict(n=1_000_000, d=784, k=30, metric="correlation", name="1Mx784 float32 correlation k=30 (MNIST-shape latent-16 synthetic)"),
    "cfg1": dict(n=1000, d=16, k=10, metric="euclidean", name="1kx16 float32 euclidean k=10 (plumbing)"),
    "small": dict(n=100_000, d=128, k=15, metric="euclidean", name="100kx128 float32 euclidean k=15 (latent-16 synthetic)"),
}
DEFAULT_WORKLOAD = "cfg4"


def latent(n, d, seed=0, latent_dim=16):
    """SURVEY.md 8d 'latent-16' generator (iid Gaussian data gives the reference recall 0.12)."""
    rs = np.random.RandomState(seed)
    Z = rs.standard_normal((n, latent_dim)).astype(np.float32)
    A = rs.standard_normal((latent_dim, d)).astype(np.float32) / np.sqrt(latent_dim)
    X = Z @ A
    X += 0.05 * rs.standard_normal((n, d)).astype(np.float32)
    return X.astype(np.float32)


def derived(wl):
    """The reference's own defaults (pynndescent_.py:1009-1012, 1135-1138; rp_trees.py:2845-2846)."""
    n, k = wl["n"], wl["k"]
    return dict(n_trees=max(3, min(12, int(round(2.0 * np.log10(n))))), n_iters=max(5, int(round(np.log2(n)))),
                leaf_size=max(60, min(256, 5 * k)), max_candidates=min(60, k), T=8)


def config_of(wl):
    """Identical for both arms (`--impl ours` / `--impl reference`): the workload, nothing about how it ran."""
    dv = derived(wl)
    ld = (wl["d"] + 3) // 4 * 4
    return {"workload": wl["name"], "n": wl["n"], "d": wl["d"], "k": wl["k"], "metric": wl["metric"], "n_trees": dv["n_trees"],
            "leaf_size": dv["leaf_size"], "max_candidates": dv["max_candidates"], "emulated_numba_threads": dv["T"],
            "l2": "inputs (X = %.0f MB) exceed the 126 MB L2; no explicit flush" % (wl["n"] * ld * 4 / 1e6)}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", os.environ.get("NND_BENCH_SMI_MS", "200")],
                                         stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        # nvidia-smi takes several hundred ms to initialise NVML and holds driver locks while it does: wait for its first
        # sample so that the start-up does not land inside the timed region (the steady 5 Hz polling does)
        t0 = time.time()
        while self.proc is not None and not self.rows and time.time() - t0 < 5.0:
            time.sleep(0.02)

    def _read(self):
        for line in self.proc.stdout:
            p = [x.strip() for x in line.split(",")]
            if len(p) >= 6:
                self.rows.append(p)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [int(r[0]) for r in self.rows if r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def exact_knn_gpu(X, rows, k, metric):
    """Exact kNN of X[rows] (brute force in float64 with torch on the GPU; measurement utility, not part of the timed path)."""
    import torch
    Xd = torch.from_numpy(X).cuda().double()
    if metric == "correlation":     # correlation distance = cosine distance of the row-centred vectors
        Xd = Xd - Xd.mean(dim=1, keepdim=True)
    if metric in ("cosine", "correlation"):
        Xd = Xd / Xd.norm(dim=1, keepdim=True).clamp_min(1e-30)
    out = []
    n2 = None if metric in ("cosine", "correlation") else (Xd * Xd).sum(1)
    for c0 in range(0, len(rows), 100):
        r = torch.from_numpy(rows[c0:c0 + 100]).cuda()
        q = Xd[r]
        dm = (1.0 - q @ Xd.T) if metric in ("cosine", "correlation") else ((q * q).sum(1)[:, None] + n2[None, :] - 2.0 * q @ Xd.T)
        out.append(dm.topk(k, dim=1, largest=False).indices.cpu().numpy())
        del dm
    del Xd
    return np.concatenate(out)


def recall_rows(n, n_sample=1000, seed=0):
    return np.random.RandomState(seed).choice(n, min(n_sample, n), replace=False)


def recall_at_k(X, graph_idx, k, metric, n_sample=1000, seed=0, true=None):
    """recall@k of `graph_idx` against exact kNN on the fixed row sample (the fixture's stored exact neighbours if given)."""
    rows = recall_rows(X.shape[0], n_sample, seed)
    if true is None:
        true = exact_knn_gpu(X, rows, k, metric)
    hits = sum(len(np.intersect1d(true[i], graph_idx[r])) for i, r in enumerate(rows))
    return hits / (len(rows) * k)


def load_fixture(workload):
    p = os.path.join(ROOT, "tests", "golden", workload + "_full_oracle.json")
    return json.load(open(p)) if os.path.exists(p) else None


def cpu_threads():
    """Threads for the CPU port.  The reference's candidate build and update application scan ALL edges / updates once
    per thread (utils.py:261-306, 709-731), so its wall time stops improving -- and then degrades -- beyond a handful of
    threads (thread sweep on the 128-core bench host: profiles/cpu_thread_sweep_r02.json).  16 is the best case for the CPU."""
    return max(1, min(os.cpu_count() or 1, int(os.environ.get("NND_BENCH_CPU_THREADS", "16"))))


def run_cpu_port(X, metric, k, seed, threads):
    """One build with the oracle port on the host cores (the only place bench.py executes oracle/)."""
    from oracle import nnd_oracle as orc
    orc.lib().orc_set_num_threads(threads)
    t0 = time.perf_counter()
    o = orc.OracleNNDescent(X, metric, n_neighbors=k, random_state=seed, n_threads=threads, forest="dfs")
    dt = time.perf_counter() - t0
    return dt, o


def full_size_cpu_note(workload):
    """A one-off full-size CPU build recorded under profiles/ (too long for every bench run), if there is one."""
    p = os.path.join(ROOT, "profiles", "cpu_full_%s_r02.json" % workload)
    return json.load(open(p)) if os.path.exists(p) else None


def reference_arm(args, wl):
    """The reference's CPU implementation of the path on this box's host cores (oracle port of the numba path; rank 0 only).
    Same `config` as the GPU arm; every step is ONE full build of a bounded sample (the first `--ref-sample` points of the
    workload's matrix) so that K + W steps end within a few minutes -- a full 10M-point CPU build takes minutes by itself
    (one is recorded in profiles/cpu_full_cfg4_r02.json and quoted in `cpu_baseline.full_size`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = cpu_threads()
    n_s = min(wl["n"], args.ref_sample)
    X = latent_prefix(wl, n_s)
    times = []
    for i in range(args.warmup + args.steps):
        dt, o = run_cpu_port(X, wl["metric"], wl["k"], 42, cores)
        if i >= args.warmup:
            times.append(dt)
    t = float(np.mean(times))
    val = n_s * wl["k"] / t
    sample = (f"first {n_s} points of the workload's matrix, one full build per step (oracle port of the numba path, OpenMP, "
              f"T={cores} of {os.cpu_count()} host cores)")
    cb = {"value": val, "unit": "edges/s", "cores": cores, "kind": "port", "sample": sample}
    fs = full_size_cpu_note(args.workload)
    if fs:
        cb["full_size"] = fs
    out = {"impl": "reference", "metric": "kNN-graph build edges/sec", "value": val, "unit": "edges/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic latent-16 (SURVEY 8d), seed 0, index random_state 42",
           "config": config_of(wl), "cpu_baseline": cb,
           "e2e": {"value": val, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def reference_thread_sweep(args, wl):
    """`--impl reference --thread-sweep`: the CPU port's build time against its OpenMP thread count on the first --ref-sample points
    (what justifies cpu_threads(); profiles/cpu_thread_sweep_r02.json)."""
    n_s = min(wl["n"], args.ref_sample)
    X = latent_prefix(wl, n_s)
    rows = []
    for T in (4, 8, 16, 32, 64, 128):
        if T > (os.cpu_count() or 1):
            break
        dt, o = run_cpu_port(X, wl["metric"], wl["k"], 42, T)
        rows.append({"threads": T, "seconds": round(dt, 3), "edges_per_s": n_s * wl["k"] / dt, "iterations": int(o.n_iters_run)})
        print(json.dumps(rows[-1]), file=sys.stderr, flush=True)
    print(json.dumps({"impl": "reference", "mode": "thread-sweep", "n": n_s, "d": wl["d"], "k": wl["k"], "host_cores": os.cpu_count(), "rows": rows}), flush=True)


def reference_full_size(args, wl):
    """`--impl reference --full-size`: ONE build of the whole workload by the CPU port (minutes at 10M points;
    profiles/cpu_full_cfg4_r02.json), with its recall on the fixture's rows."""
    cores = cpu_threads()
    X = latent(wl["n"], wl["d"], 0)
    dt, o = run_cpu_port(X, wl["metric"], wl["k"], 42, cores)
    gi, _ = o.neighbor_graph
    out = {"impl": "reference", "mode": "full-size", "workload": wl["name"], "n": wl["n"], "seconds": round(dt, 1), "edges_per_s": wl["n"] * wl["k"] / dt,
           "threads": cores, "host_cores": os.cpu_count(), "iterations": int(o.n_iters_run), "kind": "port (oracle/nnd_oracle.c, dfs forest)"}
    fx = load_fixture(args.workload)
    if fx and "exact_knn_of_sample" in fx:
        true = np.asarray(fx["exact_knn_of_sample"]); rr = recall_rows(wl["n"])
        out["recall_at_k"] = float(np.mean([len(np.intersect1d(true[i], gi[r])) for i, r in enumerate(rr)]) / wl["k"])
    print(json.dumps(out), flush=True)


def latent_prefix(wl, n_s):
    """The first n_s rows of latent(n, d) without materialising all n rows twice: the generator draws Z (n x 16) first, then A,
    then the noise row by row, so the prefix of the noise stream is the prefix of the matrix."""
    n, d = wl["n"], wl["d"]
    rs = np.random.RandomState(0)
    Z = rs.standard_normal((n, 16)).astype(np.float32)[:n_s]
    A = rs.standard_normal((16, d)).astype(np.float32) / np.sqrt(16)
    X = Z @ A
    X += 0.05 * rs.standard_normal((n_s, d)).astype(np.float32)
    return X.astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=list(WORKLOADS))
    ap.add_argument("--ref-sample", type=int, default=300_000)
    ap.add_argument("--cpu-sample", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--thread-sweep", action="store_true", help="with --impl reference: CPU port build time vs thread count")
    ap.add_argument("--full-size", action="store_true", help="with --impl reference: one full-size CPU build of the workload")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) == 0 and args.thread_sweep:
            return reference_thread_sweep(args, wl)
        if int(os.environ.get("RANK", "0")) == 0 and args.full_size:
            return reference_full_size(args, wl)
        return reference_arm(args, wl)

    import torch
    import torch.distributed as dist
    from pynndescent_b200 import NNDescent
    from pynndescent_b200 import distances as pdist
    from pynndescent_b200.engine import Engine

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    n, d, k, metric = wl["n"], wl["d"], wl["k"], wl["metric"]
    dv = derived(wl)
    n_trees, n_iters, leaf_size, C_, T = dv["n_trees"], dv["n_iters"], dv["leaf_size"], dv["max_candidates"], dv["T"]
    X = latent(n, d, seed=0)
    metric_id, dev_corr, host_corr, angular = pdist.resolve(metric)

    eng = Engine(X, metric_id, k)
    stream = eng.stream
    sb = None
    if world > 1:
        from pynndescent_b200.sharded import ShardedBuilder
        sb = ShardedBuilder(eng)

    def seeds():
        rs = np.random.RandomState(42)
        rng = rs.randint(-2 ** 31 + 1, 2 ** 31 - 1, 3).astype(np.int64); rs.randint(-2 ** 31 + 1, 2 ** 31 - 1, 3)
        trees = rs.randint(-2 ** 31 + 1, 2 ** 31 - 1, size=(n_trees, 3)).astype(np.int64)
        return rng, trees

    def one_build(timings=None):
        rng, trees = seeds()
        if sb is not None:
            i_t, d_t, it_, ch_ = sb.build(rng, trees, leaf_size, angular, C_, T, n_iters, 0.001, dev_corr, timings=timings)
            return i_t, d_t, it_, ch_, [0, 0, 0, 0]
        return eng.build(rng, trees, leaf_size, angular, C_, T, n_iters, 0.001, dev_corr, timings=timings)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_build()
    eng.profile(reset=True)
    launches0 = eng.launch_count()
    sampler = ClockSampler(local); sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        idx_t, dist_t, iters, c_hist, stats = one_build()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    prof = eng.profile(reset=True)
    launches = eng.launch_count() - launches0
    gi = idx_t.cpu().numpy(); gd = dist_t.cpu().numpy()
    if host_corr is not None:          # what NNDescent.neighbor_graph returns (cosine / dot: 1 - 2^-d in float64, distances.py:704-711)
        gd = host_corr(gd)

    # ---- one more build OUTSIDE the timed region with the library's stage timers on (CUDA events at every stage
    # boundary) and host-side stage laps: the stage table of the JSON line
    stage_host = {}
    eng.stage_profile(True); eng.stage_ms(reset=True)
    one_build(stage_host)
    torch.cuda.synchronize()
    stage_dev = eng.stage_ms(reset=True)
    eng.stage_profile(False)

    # ---- e2e: the public API with host buffers (H2D of X and D2H of the graph inside the timed region)
    comm_bytes = sb.comm_bytes if sb is not None else 0
    comm_builds = args.steps + args.warmup + 1
    del eng, idx_t, dist_t, sb
    torch.cuda.empty_cache()
    e2e_ms = float("nan")
    if not args.no_e2e:
        Xp = torch.from_numpy(X).pin_memory().numpy()
        NNDescent(Xp, metric, n_neighbors=k, random_state=42, sharded=(world > 1)).neighbor_graph  # warm allocators (device + pinned staging)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2e_steps = max(1, min(args.steps, 3))
        g = None
        e0.record(stream)
        for _ in range(e2e_steps):
            t_a = time.perf_counter()
            idx = NNDescent(Xp, metric, n_neighbors=k, random_state=42, sharded=(world > 1))
            t_b = time.perf_counter()
            if rank == 0 or world == 1:      # every rank holds the finished graph; the user reads it once, from rank 0
                g = idx.neighbor_graph
            if rank == 0 and os.environ.get("NND_BENCH_VERBOSE"):
                print("e2e step: ctor %.1f ms, neighbor_graph %.1f ms, timings %s" % (
                    (t_b - t_a) * 1e3, (time.perf_counter() - t_b) * 1e3, {kk: round(vv * 1e3, 1) for kk, vv in idx.timings.items()}),
                    file=sys.stderr, flush=True)
            del idx          # one index alive at a time: the next build reuses its (cached) device and pinned blocks
            g = None         # ... and the graph copy's pinned staging goes back to torch's host allocator cache
        e1.record(stream)
        barrier()
        e2e_ms = e0.elapsed_time(e1) / e2e_steps
        del Xp

    t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_step = ms / args.steps
    value = n * k / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel (k_join_batch, the local join), DESIGN.md section 7:
    # algorithmic bytes = 4*ld*G (gathered rows) + 4*G (thresholds) + 8*N*C (candidate lists, per iteration)
    #                     + 12*U (half-updates written), time = CUDA-event duration of those launches
    peak, peak_src = measured_peaks()
    ld = (d + 3) // 4 * 4
    jj = prof["join"]
    iters_total = iters * args.steps
    bytes_alg = 4.0 * ld * jj["rows"] + 4.0 * jj["rows"] + 8.0 * (n / world) * C_ * iters_total + 12.0 * jj["pushes"]
    ach = bytes_alg / (jj["ms"] * 1e-3) / 1e9 if jj["ms"] > 0 else 0.0
    lf = prof["leaf"]
    leaf_bytes = 4.0 * ld * lf["rows"] + 4.0 * lf["rows"] + 12.0 * lf["pushes"]
    leaf_ach = leaf_bytes / (lf["ms"] * 1e-3) / 1e9 if lf["ms"] > 0 else 0.0

    # ---- parity + recall against the full-size oracle fixture of this workload
    fx = load_fixture(args.workload)
    true = np.asarray(fx["exact_knn_of_sample"]) if fx and "exact_knn_of_sample" in fx else None
    rec = recall_at_k(X, gi, k, metric, true=true)
    parity = {"idx_sha256": hashlib.sha256(np.ascontiguousarray(gi, np.int32).tobytes()).hexdigest(),
              "dist_sha256": hashlib.sha256(np.ascontiguousarray(gd).tobytes()).hexdigest(),
              "iterations": int(iters), "c_hist": [int(c) for c in c_hist], "equals_cpu_oracle_graph": None,
              "note": "identical for every --gpus N by construction (emission-sequence replay); the oracle fixture is the level-forest "
                      "CPU oracle's graph of the same workload (tests/golden/make_full_size_checksum.py)"}
    if fx and "idx_sha256" in fx:
        parity["equals_cpu_oracle_graph"] = bool(fx["idx_sha256"] == parity["idx_sha256"] and fx["dist_sha256"] == parity["dist_sha256"]
                                                 and fx["c_hist"] == parity["c_hist"])
    recall = {"at_k": rec, "sample": "1000 rows, np.random.RandomState(0).choice(n, 1000, replace=False)",
              "exact_knn": "fixture" if true is not None else "float64 brute force on the GPU",
              "reference_dfs_oracle_same_rows": fx.get("recall_dfs_oracle") if fx else None,
              "level_oracle_same_rows": fx.get("recall_level_oracle") if fx else None}
    if recall["reference_dfs_oracle_same_rows"] is not None:
        recall["delta_vs_reference_pt"] = 100.0 * (rec - recall["reference_dfs_oracle_same_rows"])

    traffic, traffic_src = None, "not captured for this workload"
    traffic_file = os.path.join(ROOT, "profiles", "k_join_traffic.json")
    if os.path.exists(traffic_file):
        tj = json.load(open(traffic_file))
        ent = tj.get(args.workload) if isinstance(tj.get(args.workload), dict) else None
        if ent:
            traffic = ent.get("dram_bytes_per_launch"); traffic_src = "static: " + ent.get("source", "profiles/k_join_traffic.json")
    out = {
        "metric": "kNN-graph build edges/sec", "value": value, "unit": "edges/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic latent-16 (SURVEY 8d), seed 0, index random_state 42",
        "config": config_of(wl),
        "run": {"iterations": int(iters),
                "multi_gpu": ("one %d-point graph over %d GPUs: heap rows sharded by range (utils.py:259-264), X replicated; "
                              "%.1f MB/build/rank exchanged over NCCL" % (n, world, comm_bytes / 1e6 / max(comm_builds, 1)))
                if world > 1 else "single GPU",
                "schedule_switches": {kk: vv for kk, vv in sorted(os.environ.items()) if kk.startswith("NND_") and not kk.startswith("NND_BENCH")}},
        "recall_at_k": rec, "recall": recall, "parity": parity,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": "k_join_batch (local join, utils.py:536-658)", "bound": "hbm", "achieved": ach, "peak": peak,
                     "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "kernel_ms_per_step": jj["ms"] / args.steps, "launches_per_step": jj["launches"] / args.steps,
                     "rows_gathered_per_step": jj["rows"] / args.steps, "pairs_per_step": jj["pairs"] / args.steps,
                     "leaf_join": {"achieved": leaf_ach, "frac": leaf_ach / peak, "kernel_ms_per_step": lf["ms"] / args.steps,
                                   "pairs_per_step": lf["pairs"] / args.steps}},
        "stage_ms": {"device_events": {kk: round(vv, 3) for kk, vv in stage_dev.items()},
                     "host_laps": {kk: round(vv * 1e3, 3) for kk, vv in stage_host.items()},
                     "note": "one extra build outside the timed region, rank 0"},
    }
    if not args.no_e2e:
        out["e2e"] = {"value": n * k / (e2e_ms * 1e-3), "unit": "edges/s", "ms_per_step": e2e_ms,
                      "h2d_bytes_per_step": int(n * d * 4), "d2h_bytes_per_step": int(n * k * 8),
                      "note": "NNDescent(X_pinned_host) + .neighbor_graph per step; X is replicated on the GPUs: every rank uploads 1 / n_gpus of "
                              "its rows over PCIe and an NVLink all-gather completes the copies; the graph is read back once (rank 0)"}
    if not args.no_cpu_baseline:
        cores = cpu_threads()
        n_s = min(n, args.cpu_sample)
        dt, o = run_cpu_port(X[:n_s].copy(), metric, k, 42, cores)
        out["cpu_baseline"] = {"value": n_s * k / dt, "unit": "edges/s", "cores": cores, "kind": "port",
                               "sample": f"one full build of the first {n_s} points of the same matrix by the oracle "
                                         f"port of the numba path (OpenMP, T={cores} of {os.cpu_count()} host cores), {dt:.1f} s"}
        fs = full_size_cpu_note(args.workload)
        if fs:
            out["cpu_baseline"]["full_size"] = fs
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
