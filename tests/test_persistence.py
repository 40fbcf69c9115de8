"""an index in this package's format, pickled in the REFERENCE's __getstate__ layout, is loaded by the stock
reference with pickle.load and answers queries identically.  CPU-only; needs the reference tree (build container) and is
skipped where it is absent (the GPU box)."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden

LOADER = r'''
import pickle, sys, numpy as np
sys.path.insert(0, sys.argv[1])
from oracle.ref_shim import import_reference
pyn = import_reference()
obj = pickle.load(open(sys.argv[2], "rb"))
assert type(obj) is pyn.NNDescent, type(obj)
G = np.load(sys.argv[3])
qi, qd = obj.query(G["Q"], k=int(G["k_query"]), epsilon=float(G["epsilon"]))
qi2, qd2 = obj.query(G["Q"], k=5, epsilon=0.3)
gi, gd = obj.neighbor_graph
# update() reads _input_dtype / _angular_trees / n_trees_after_update from the unpickled state (pynndescent_.py:2407-2535)
n0 = gi.shape[0]
obj.update(xs_fresh=np.ascontiguousarray(G["Q"][:25]))
ui, ud = obj.neighbor_graph
assert ui.shape[0] == n0 + 25 and (ui[:, 0] >= 0).all()
np.savez(sys.argv[4], qi=qi, qd=qd, qi2=qi2, qd2=qd2, gi=gi, gd=gd)
'''


def _fake_prepared_index(G):
    """A pynndescent_b200.NNDescent as it looks after a GPU build + prepare(), filled from the golden (the structures in
    the golden are the reference's own, and the GPU's prepare() is tested equal to them in tests/test_gpu_query.py)."""
    from pynndescent_b200 import NNDescent
    from pynndescent_b200 import distances
    metric = str(G["metric"])
    o = NNDescent.__new__(NNDescent)
    o.n_trees, o.n_neighbors, o.metric, o.metric_kwds, o.leaf_size = 7, int(G["n_neighbors"]), metric, None, None
    o.pruning_degree_multiplier, o.max_candidates, o.max_rptree_depth, o.n_iters, o.delta = 1.5, None, 200, 12, 0.001
    o.n_jobs, o.low_memory, o.compressed, o.verbose, o.random_state, o.dim = None, True, False, False, 42, G["X"].shape[1]
    o.tree_init = True
    _, _, host_corr, angular = distances.resolve(metric)
    o._angular_trees, o._distance_correction = angular, host_corr
    o._raw_data = G["X"]
    o.rng_state, o.search_rng_state = G["rng_before_prepare"], G["search_rng_state"]
    corrected = np.sqrt(G["ng_dist"]) if metric in ("euclidean", "l2") else G["ng_dist"]
    o._neighbor_graph, o._raw_graph_dist = (G["ng_idx"], corrected), G["ng_dist"]
    o._search_structures = {k: G[k] for k in ("indptr", "indices", "hyperplanes", "offsets", "children", "tree_indices",
                                              "vertex_order", "min_distance")}
    return o


REF_CASES = ["query_euclid_3k32", "query_cosine_2k24", "query_manhattan_2k20"]          # the last has no fast alternative


@pytest.fixture(scope="module")
def reference_loader_runs(tmp_path_factory):
    """One stock-reference process per case, all started together (each spends most of its time in numba's JIT for its own
    metric, so running them side by side keeps the CPU suite short); the test of a case waits for its process."""
    from oracle import ref_shim
    if not ref_shim.available():
        return None
    runs = {}
    for name in REF_CASES:
        tmp = tmp_path_factory.mktemp(name)
        G = load_golden(name)
        idx = _fake_prepared_index(G)
        state = idx.reference_state()
        assert np.array_equal(state["_raw_data"], G["data"])                 # leaf-order data (pynndescent_.py:1639-1641)
        assert state["_search_graph"].dtype == np.uint8 and state["_search_graph"].nnz == G["indices"].shape[0]
        assert "pynndescent" not in sys.modules                              # written WITHOUT the reference importable
        pk = tmp / "index.pkl"
        idx.save_for_reference(str(pk))
        assert "pynndescent" not in sys.modules and "pynndescent.pynndescent_" not in sys.modules
        script = tmp / "loader.py"; script.write_text(LOADER)
        out = tmp / "out.npz"
        proc = subprocess.Popen([sys.executable, str(script), ROOT, str(pk), os.path.join(ROOT, "tests", "golden", name + ".npz"), str(out)],
                                stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        runs[name] = (proc, out, G)
    yield runs
    for proc, _, _ in runs.values():
        if proc.poll() is None:
            proc.kill()


@pytest.mark.parametrize("name", REF_CASES)
def test_reference_loads_and_queries_our_pickle(reference_loader_runs, name):
    if reference_loader_runs is None:
        pytest.skip("reference tree not present on this machine")
    proc, out, G = reference_loader_runs[name]
    _, err = proc.communicate(timeout=900)
    assert proc.returncode == 0, err[-3000:]
    R = np.load(out)
    # the stock reference, running ITS OWN numba search on the unpickled structures, returns its golden answers
    assert np.array_equal(R["qi"], G["out_idx"]) and np.array_equal(R["qd"], G["out_dist"])
    assert np.array_equal(R["qi2"], G["out_idx_k5_e03"]) and np.array_equal(R["qd2"], G["out_dist_k5_e03"])
    assert np.array_equal(R["gi"], G["ng_idx"])


@pytest.mark.gpu
def test_reference_state_of_a_gpu_built_index():
    """The same export from a real GPU build: raw (uncorrected) graph distances, leaf-ordered data, one flat search tree."""
    import io
    from conftest import latent
    from pynndescent_b200 import NNDescent
    X = latent(3000, 24, 9)
    idx = NNDescent(X, "euclidean", n_neighbors=12, random_state=3)
    st = idx.reference_state()
    gi, gd = idx.neighbor_graph
    assert np.array_equal(st["_neighbor_graph"][0], gi)
    np.testing.assert_allclose(np.sqrt(st["_neighbor_graph"][1]), gd, rtol=1e-6)      # stored raw = squared distances
    vo = st["_vertex_order"]
    assert sorted(vo.tolist()) == list(range(3000)) and np.array_equal(st["_raw_data"], X[vo])
    hyper, offs, children, tidx, leaf_size = st["_search_forest"][0]
    assert hyper.shape == (children.shape[0], 24) and offs.shape[0] == children.shape[0] and leaf_size > 0
    assert st["_search_graph"].shape == (3000, 3000) and st["_search_graph"].nnz == idx._search_structures["indices"].shape[0]
    buf = io.BytesIO()
    idx.save_for_reference(buf)
    assert buf.tell() > X.nbytes
