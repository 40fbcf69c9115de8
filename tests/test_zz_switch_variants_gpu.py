"""Every schedule switch that is still in the library (the defaults are the winners of the round-2 hardware A/B,
the other value of each switch is the path it replaced or a fallback) must leave the result untouched: one small build per variant in a
FRESH process (the switches are read from the environment by nnd_create / the engine), indices + distances + per-iteration accepted
pushes identical to the default build.  The same harness (scripts/r2_ab.py) produced profiles/ab_r02_*.json at 1M and 10M points.

Sorts after every other GPU test file on purpose: the parity tests proper report first.
"""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_every_switch_variant_builds_the_same_graph(tmp_path):
    import importlib.util
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    spec = importlib.util.spec_from_file_location("r2_ab", os.path.join(ROOT, "scripts", "r2_ab.py"))
    ab = importlib.util.module_from_spec(spec); spec.loader.exec_module(ab)
    n, d, k = 60000, 32, 15
    results = {}
    for name, switches in ab.CONFIGS.items():
        env = {kk: vv for kk, vv in os.environ.items() if not kk.startswith("NND_")}
        env.update(switches)
        out = tmp_path / (name.replace("+", "_") + ".json")
        p = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "r2_ab.py"), "child", str(n), str(d), str(k), "euclidean", "0", str(out)],
                           env=env, capture_output=True, text=True, timeout=300)
        assert p.returncode == 0 and out.exists(), (name, p.stderr[-2000:])
        results[name] = json.load(open(out))
    base = results["default"]
    assert base["iters"] >= 3 and len(base["c_hist"]) == base["iters"]
    for name, r in results.items():
        assert (r["idx_sha"], r["dist_sha"], r["c_hist"]) == (base["idx_sha"], base["dist_sha"], base["c_hist"]), name
