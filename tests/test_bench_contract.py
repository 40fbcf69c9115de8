"""bench.py's reference arm (`--impl reference`: the CPU implementation of the path on the host cores) runs without a GPU: its JSON
line carries the keys the driver reads, rank 0 alone prints it."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1", "--steps", "2",
                          "--warmup", "1", "--gpus", "1"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return [ln for ln in out.stdout.splitlines() if ln.startswith("{")]


def test_reference_arm_json_line():
    lines = _run()
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "kNN-graph build edges/sec" and j["unit"] == "edges/s"
    assert j["higher_is_better"] is True and j["steps"] == 2 and j["warmup"] == 1 and j["n_gpus"] == 1 and j["vs_baseline"] is None
    assert j["value"] > 0 and abs(j["value"] - 1000 * 10 / (j["ms_per_step"] * 1e-3)) < 1e-6 * j["value"]
    assert j["config"]["workload"].startswith("1kx16") and j["dtype"] == "f32" and "synthetic" in j["data"]
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_print_nothing():
    assert _run({"RANK": "1", "WORLD_SIZE": "2"}) == []


def test_both_arms_describe_the_same_config():
    """The driver compares the two arms' lines: `config` must be the workload's description, identical in both (bench.config_of)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
    j = json.loads(_run()[0])
    assert j["config"] == bench.config_of(bench.WORKLOADS["cfg1"]) and j["scaling"] == "strong"
    assert bench.DEFAULT_WORKLOAD == "cfg4" and bench.WORKLOADS["cfg4"]["n"] == 10_000_000      # the metric's own configuration
    for wl in bench.WORKLOADS.values():
        assert set(bench.config_of(wl)) >= {"workload", "n", "d", "k", "metric"}


def test_reference_sample_is_a_prefix_of_the_workload_matrix():
    """The reference arm builds the FIRST rows of the GPU arm's matrix without generating all of it (bench.latent_prefix)."""
    import importlib.util
    import numpy as np
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
    wl = dict(n=5000, d=24)
    assert np.array_equal(bench.latent(5000, 24, 0)[:1234], bench.latent_prefix(wl, 1234))
