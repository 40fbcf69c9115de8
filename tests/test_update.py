"""NNDescent.update() (pynndescent_.py:2381-2553).

Goldens (tests/golden/make_golden_update.py) come from the REAL reference: index built on xs_orig, then
update(xs_fresh, xs_updated, updated_indices), with the leaf_array of both forests captured.  Given those two forests the
whole path is deterministic, so the final graph must equal the reference's:

* CPU: the oracle's restatement of update() reproduces the reference (indices exact, raw distances <= 1e-5 relative);
* CPU: the PRODUCT's host logic (validation, row replacement, invalidation of touched entries, stage order, tree count
  bookkeeping) reproduces it too when its stages are executed by the oracle-backed stand-in engine (tests/oracle_engine.py);
* GPU: the product end to end on the kernels.
"""
import numpy as np
import pytest

from conftest import load_golden

CASES = {"update_fresh_euclid_1200x16": "euclidean", "update_changed_euclid_1200x16": "euclidean",
         "update_changed_cosine_1000x20": "cosine", "update_only_changed_euclid_800x12": "euclidean"}


def _args(g):
    ui = g["updated_indices"]
    return dict(xs_fresh=g["xs_fresh"] if len(g["xs_fresh"]) else None, xs_updated=g["xs_updated"] if len(ui) else None,
                updated_indices=ui if len(ui) else None)


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_update_matches_reference(name):
    from oracle import nnd_oracle as orc
    g = load_golden(name); metric = CASES[name]
    o = orc.OracleNNDescent(g["xs_orig"], metric, n_neighbors=int(g["k"]), random_state=int(g["seed"]), n_threads=int(g["T"]),
                            leaf_array=g["leaf_array_build"])
    assert np.array_equal(o._neighbor_graph[0], g["graph0_idx"])
    assert np.array_equal(o.rng_state, g["rng_state_after_build"])
    o.update(leaf_array=g["leaf_array_update"], **_args(g))
    assert o.n_trees == int(g["n_trees_after"])
    assert np.array_equal(o._neighbor_graph[0], g["graph1_idx"])
    np.testing.assert_allclose(o._neighbor_graph[1], g["graph1_raw_dist"], rtol=1e-5)
    assert np.array_equal(o.rng_state, g["rng_state_after_update"])
    if "graph2_idx" in g.files:          # a second update keeps the reduced tree count (reference test :646-653)
        o.update(xs_fresh=g["xs_fresh2"], leaf_array=g["leaf_array_update2"])
        assert o.n_trees == int(g["n_trees_after2"])
        assert np.array_equal(o._neighbor_graph[0], g["graph2_idx"])


def _product_roundtrip(g, metric):
    from pynndescent_b200 import NNDescent
    idx = NNDescent(g["xs_orig"], metric, n_neighbors=int(g["k"]), random_state=int(g["seed"]), n_jobs=int(g["T"]),
                    leaf_array=g["leaf_array_build"])
    gi0, _ = idx.neighbor_graph
    assert np.array_equal(gi0, g["graph0_idx"])
    n_trees0 = idx.n_trees
    idx.update(leaf_array=g["leaf_array_update"], **_args(g))
    gi, gd = idx.neighbor_graph
    assert idx.n_trees == int(g["n_trees_after"]) and idx.n_trees_after_update == max(2, int(np.round(n_trees0 / 3)))
    assert idx._raw_data.shape[0] == g["graph1_idx"].shape[0]
    assert np.array_equal(gi, g["graph1_idx"])
    np.testing.assert_allclose(np.asarray(gd, np.float64), np.asarray(g["graph1_dist"], np.float64), rtol=1e-5, atol=1e-7)
    if "graph2_idx" in g.files:
        idx.update(xs_fresh=g["xs_fresh2"], leaf_array=g["leaf_array_update2"])
        assert idx.n_trees == int(g["n_trees_after2"])
        assert np.array_equal(idx.neighbor_graph[0], g["graph2_idx"])
    return idx


@pytest.mark.parametrize("name", list(CASES))
def test_product_update_host_logic_on_oracle_engine(name, monkeypatch):
    """The product's update() with its stages executed by the oracle (no GPU): checks everything the host does."""
    import pynndescent_b200.pynndescent_ as mod
    from oracle_engine import OracleEngine
    monkeypatch.setattr(mod, "Engine", OracleEngine)
    idx = _product_roundtrip(load_golden(name), CASES[name])
    assert idx._engine.calls[:3] == ["init_from_graph", "leaf_init_from_array", "descent"]
    assert not idx._engine.heap_flag.numpy()[idx._engine.heap_idx.numpy() < 0].any()


def test_update_argument_errors(monkeypatch):
    """pynndescent_.py:2413-2441."""
    import pynndescent_b200.pynndescent_ as mod
    from oracle_engine import OracleEngine
    from pynndescent_b200 import NNDescent
    monkeypatch.setattr(mod, "Engine", OracleEngine)
    g = load_golden("update_only_changed_euclid_800x12")
    idx = NNDescent(g["xs_orig"], "euclidean", n_neighbors=10, random_state=1, leaf_array=g["leaf_array_build"])
    with pytest.raises(ValueError, match="updated_indices must also be provided"):
        idx.update(xs_updated=g["xs_updated"])
    with pytest.raises(ValueError, match="must match"):
        idx.update(xs_updated=g["xs_updated"], updated_indices=[1, 2, 3])
    with pytest.raises(ValueError, match="Could not convert"):
        idx.update(xs_updated=g["xs_updated"][:1], updated_indices=["a"])
    bad = g["xs_updated"][:1].copy(); bad[0, 0] = np.nan
    before = idx.neighbor_graph[0].copy()
    with pytest.raises(ValueError, match="NaN"):
        idx.update(xs_fresh=bad)
    assert np.array_equal(idx.neighbor_graph[0], before) and idx._raw_data.shape[0] == 800   # a rejected update changes nothing
    with pytest.warns(UserWarning, match="will be ignored"):
        idx.update(updated_indices=[1, 2], leaf_array=g["leaf_array_update"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_product_update_matches_reference_gpu(name):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    _product_roundtrip(load_golden(name), CASES[name])


@pytest.mark.gpu
def test_update_with_gpu_forest_recall_and_requery():
    """No injected forest: update() grows its own (level-synchronous) forest; quality check against brute force, and a
    prepared index is re-prepared so that query() answers on the new data (pynndescent_.py:2539-2553)."""
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from conftest import latent
    from pynndescent_b200 import NNDescent
    X = latent(24000, 32, 5)
    idx = NNDescent(X[:20000], "euclidean", n_neighbors=15, random_state=3)
    idx.prepare()
    repl = np.arange(0, 2000, 2)
    idx.update(xs_fresh=X[22000:], xs_updated=X[20000:21000], updated_indices=repl)
    Y = X[:20000].copy(); Y[repl] = X[20000:21000]; Y = np.vstack([Y, X[22000:]])
    gi, gd = idx.neighbor_graph
    assert gi.shape == (22000, 15)
    rows = np.random.RandomState(0).choice(22000, 500, replace=False)
    d2 = (Y[rows].astype(np.float64) ** 2).sum(1)[:, None] + (Y.astype(np.float64) ** 2).sum(1)[None, :] \
        - 2.0 * Y[rows].astype(np.float64) @ Y.astype(np.float64).T
    true = np.argsort(d2, axis=1, kind="stable")[:, :15]
    rec = np.mean([len(np.intersect1d(true[i], gi[r])) / 15 for i, r in enumerate(rows)])
    assert rec >= 0.90, rec          # the oracle's update() reaches 0.973 on this input
    exact = np.sqrt(((Y[rows, None, :].astype(np.float64) - Y[gi[rows]].astype(np.float64)) ** 2).sum(-1))
    np.testing.assert_allclose(gd[rows], exact, rtol=1e-5, atol=1e-6)
    qi, qd = idx.query(Y[rows[:50]], k=5)
    assert qi.shape == (50, 5) and np.mean(qi[:, 0] == rows[:50]) >= 0.8   # indexed points find themselves on the re-prepared index
    # the same update on the oracle (level forest = the GPU forest's scheme): bit-identical graph
    from oracle import nnd_oracle as orc
    o = orc.OracleNNDescent(X[:20000], "euclidean", n_neighbors=15, random_state=3, n_threads=8, forest="level")
    o.update(xs_fresh=X[22000:], xs_updated=X[20000:21000], updated_indices=repl)
    oi, od = o.neighbor_graph
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)
