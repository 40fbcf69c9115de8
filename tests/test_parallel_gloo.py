"""CPU-only: the N>1 host logic over the gloo backend, world_size 2 (rendezvous on 127.0.0.1)."""
import os
import socket
import subprocess
import sys

import numpy as np

from conftest import ROOT

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["REPO_ROOT"])
import importlib.util
spec = importlib.util.spec_from_file_location("parallel", os.path.join(os.environ["REPO_ROOT"], "pynndescent_b200", "parallel.py"))
par = importlib.util.module_from_spec(spec); spec.loader.exec_module(par)
import torch.distributed as dist
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
lo, hi = par.owner_ranges(1001, w)[r]
# every rank "builds" its own range; aggregate edges and the slowest time as bench.py does
edges = par.sum_over_ranks((hi - lo) * 15)
t = par.max_over_ranks(1.0 + r)
if r == 0:
    print("RESULT", int(edges), t, flush=True)
dist.destroy_process_group()
'''


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def test_owner_ranges_match_reference_rule():
    import importlib.util
    spec = importlib.util.spec_from_file_location("parallel", os.path.join(ROOT, "pynndescent_b200", "parallel.py"))
    par = importlib.util.module_from_spec(spec); spec.loader.exec_module(par)
    for n, w in ((1000, 8), (1001, 2), (17, 4), (10_000_000, 8)):
        rr = par.owner_ranges(n, w)
        assert rr[0][0] == 0 and rr[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(rr, rr[1:]))
        block = n // w + 1                                  # utils.py:259
        assert all(hi - lo <= block for lo, hi in rr)
        v = np.array([0, n // 2, n - 1])
        assert np.array_equal(par.owner_of(v, n, w), v // block)


def test_world_size_2_gloo_reductions(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, REPO_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][0].split()
    assert int(line[1]) == 1001 * 15 and float(line[2]) == 2.0
