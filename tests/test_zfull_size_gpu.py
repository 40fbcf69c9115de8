"""Full-size GPU test: BASELINE.json configs[1] (1M x 128 float32 euclidean k=15, latent-16 synthetic) through the public
API, checked by the size-independent properties of a kNN graph (the oracle needs minutes at this size):

* every row sorted by ascending distance, the point itself first at distance 0, no duplicate neighbours
  (reference tests/test_pynndescent_.py:299-369 check the same on their data);
* the reported distances are the true euclidean distances of the reported neighbours (rtol 1e-5, the bar of
  BASELINE.json's north_star);
* determinism: a second build from the same (data, random_state) is bit-identical (reference test_deterministic
  :279-292);
* recall@15 against brute force on a 1000-row sample >= the reference's 0.885 minus sampling slack;
* the whole 1,000,000 x 15 graph equals the CPU oracle's, bit for bit (sha256 of indices and of distances, per-iteration
  accepted-push counts; fixture tests/golden/cfg2_full_oracle.json written by tests/golden/make_full_size_checksum.py).

Runs last (file name) so that the small, oracle-checked parity tests report first.
"""
import numpy as np
import pytest

from conftest import latent

pytestmark = pytest.mark.gpu

# |recall_gpu - recall of the reference-scheme (dfs forest) oracle on the same rows|: north_star's 0.5 pt
RECALL_BAR = {"cfg2": 0.005, "cfg3": 0.005, "cfg4": 0.005, "cfg5": 0.005}


def test_cfg2_full_size_properties_and_determinism():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from bench import recall_at_k
    from pynndescent_b200 import NNDescent
    n, d, k = 1_000_000, 128, 15
    X = latent(n, d, 0)
    idx = NNDescent(X, "euclidean", n_neighbors=k, random_state=42)
    gi, gd = idx.neighbor_graph
    assert gi.shape == (n, k) and gi.dtype == np.int32 and gd.shape == (n, k) and gd.dtype == np.float32
    assert gi.min() >= 0 and gi.max() < n                                   # every row filled
    self_first = gi[:, 0] == np.arange(n)                                   # the (p, p) pair of the local join puts p in its own row
    assert self_first.mean() >= 0.9999 and np.all(gd[self_first, 0] == 0)
    assert np.all(np.diff(gd, axis=1) >= 0)                                 # ascending
    srt = np.sort(gi, axis=1)
    assert np.all(srt[:, 1:] != srt[:, :-1])                                # no duplicate neighbours
    rows = np.random.RandomState(1).choice(n, 4000, replace=False)
    exact = np.sqrt(((X[rows, None, :].astype(np.float64) - X[gi[rows]].astype(np.float64)) ** 2).sum(-1))
    np.testing.assert_allclose(gd[rows], exact, rtol=1e-5, atol=1e-6)
    import hashlib
    import json
    import os
    from conftest import GOLDEN
    g = json.load(open(os.path.join(GOLDEN, "cfg2_full_oracle.json")))
    # recall on the fixture's fixed 1000 rows against ITS exact neighbours (float64 brute force, make_full_size_checksum.py), next to
    # the reference-scheme (dfs forest) oracle's recall on the same rows: north_star's bar is 0.5 pt.  The two builds use different
    # random forests (the reference's rng is tied to its depth-first recursion), and recall moves by about that
    # much from seed to seed (tests/studies/recall_seed_spread_cfg2.json), so the bar is checked with that spread stated
    rec = recall_at_k(X, gi, k, "euclidean", n_sample=1000, seed=0, true=np.asarray(g["exact_knn_of_sample"]))
    assert abs(rec - recall_at_k(X, gi, k, "euclidean", n_sample=1000, seed=0)) < 1e-12       # fixture neighbours == GPU brute force
    # seed 42 alone: -0.65 pt (0.8797 vs 0.8861); the seed-to-seed standard deviation of either forest scheme is 0.16-0.28 pt, so a
    # single pair of builds cannot resolve a 0.5 pt bar -- it is asserted on the five-seed means below, and the one-seed figure is
    # bounded by the bar plus one standard deviation of the difference
    assert abs(rec - g["recall_dfs_oracle"]) <= RECALL_BAR["cfg2"] + 0.0035, (rec, g["recall_dfs_oracle"])
    iters, c_hist = idx.n_iters_run, list(idx.c_hist)
    # bit-for-bit against the CPU oracle's build of the same 1M-point workload (tests/golden/make_full_size_checksum.py):
    # per-iteration accepted-push counts, four literal rows, then sha256 of all 15M indices and of all 15M distances
    assert iters == g["iters"] and [int(c) for c in c_hist] == g["c_hist"]
    for r, row in g["rows"].items():
        assert gi[int(r)].tolist() == row["idx"] and [float(v) for v in gd[int(r)]] == row["dist"]
    assert hashlib.sha256(np.ascontiguousarray(gi, np.int32).tobytes()).hexdigest() == g["idx_sha256"]
    assert hashlib.sha256(np.ascontiguousarray(gd, np.float32).tobytes()).hexdigest() == g["dist_sha256"]
    del idx
    # five seeds against the five reference-scheme (dfs forest) oracle builds of tests/studies/recall_seed_spread.py: every GPU
    # build has exactly the recall of the level-scheme oracle build of its seed (same graph), and the MEANS differ by < 0.5 pt
    study = json.load(open(os.path.join(os.path.dirname(GOLDEN), "studies", "recall_seed_spread_cfg2.json")))
    true = np.asarray(g["exact_knn_of_sample"])
    recs = []
    for run in [r for r in study["runs"] if r["forest"] == "level"]:
        gs = NNDescent(X, "euclidean", n_neighbors=k, random_state=run["seed"]).neighbor_graph[0]
        recs.append(recall_at_k(X, gs, k, "euclidean", true=true))
        assert abs(recs[-1] - run["recall_at_15"]) < 1e-12, (run, recs[-1])
    assert abs(float(np.mean(recs)) - study["dfs"]["mean"]) <= RECALL_BAR["cfg2"], (np.mean(recs), study["dfs"])
    idx2 = NNDescent(X, "euclidean", n_neighbors=k, random_state=42)
    gi2, gd2 = idx2.neighbor_graph
    assert idx2.n_iters_run == iters and list(idx2.c_hist) == c_hist
    assert np.array_equal(gi, gi2) and np.array_equal(gd, gd2)


@pytest.mark.parametrize("name", ["cfg3", "cfg4", "cfg5"])
def test_full_size_checksum_of_the_other_configs(name):
    """BASELINE.json configs[2] (2M x 256 float32 cosine k=15: angular trees, alternative_cosine + correction), configs[3] (the
    metric's own 10M x 128 euclidean k=15 graph, here on ONE GPU) and configs[4] (1M x 784 correlation k=30: f64 accumulation, leaves
    of up to 150 points, C = 30): the whole graph equals the CPU oracle's, bit for bit (fixtures:
    tests/golden/make_full_size_checksum.py cfg3 | cfg4 | cfg5), and its recall on the fixture's rows is within the bar of the
    reference-scheme oracle's."""
    import hashlib
    import json
    import os
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from conftest import GOLDEN
    from pynndescent_b200 import NNDescent
    path = os.path.join(GOLDEN, name + "_full_oracle.json")
    if name == "cfg5" and os.environ.get("NND_FULL_CFG5") != "1":
        # 1M x 784 correlation k=30 is ~4e9 pair evaluations of 784 f64 fma each on the generic join: minutes of GPU time, and the
        # configuration has not run at full size on hardware yet -- scripts/r2_measure.sh sets NND_FULL_CFG5=1 and runs it on its own
        pytest.skip("set NND_FULL_CFG5=1 (minutes of GPU time; first run is scripted in scripts/r2_measure.sh)")
    if not os.path.exists(path):
        pytest.skip("fixture not generated (minutes of oracle time): python tests/golden/make_full_size_checksum.py " + name)
    g = json.load(open(path))
    X = latent(g["n"], g["d"], 0)
    idx = NNDescent(X, g["metric"], n_neighbors=g["k"], random_state=g["random_state"])
    gi, gd = idx.neighbor_graph
    assert gi.shape == (g["n"], g["k"]) and str(np.asarray(gd).dtype) == g["dist_dtype"]
    assert idx.n_iters_run == g["iters"] and [int(c) for c in idx.c_hist] == g["c_hist"]
    for r, row in g["rows"].items():
        assert gi[int(r)].tolist() == row["idx"] and [float(v) for v in gd[int(r)]] == row["dist"]
    assert hashlib.sha256(np.ascontiguousarray(gi, np.int32).tobytes()).hexdigest() == g["idx_sha256"]
    assert hashlib.sha256(np.ascontiguousarray(gd).tobytes()).hexdigest() == g["dist_sha256"]
    assert np.all(np.diff(gd, axis=1) >= 0) and (gi[:, 0] == np.arange(g["n"])).mean() >= 0.9999
    if "exact_knn_of_sample" in g and g.get("recall_dfs_oracle") is not None:
        from bench import recall_at_k
        rec = recall_at_k(X, gi, g["k"], g["metric"], true=np.asarray(g["exact_knn_of_sample"]))
        if g.get("recall_level_oracle") is not None:
            assert abs(rec - g["recall_level_oracle"]) < 1e-12         # same graph => same recall
        assert abs(rec - g["recall_dfs_oracle"]) <= RECALL_BAR[name], (rec, g["recall_dfs_oracle"])
