"""PyNNDescentTransformer (pynndescent_.py:2556-2863; SURVEY 8b public API row).

CPU: fit_transform() on the oracle-backed engine -- CSR layout, n_neighbors + 1 convention, explicit diagonal, parameter
plumbing (sklearn get_params / clone).  GPU: the reference's own test_transformer_equivalence (tests/test_pynndescent_.py
:234-258): transform(test) == index.query(test) of an index built with n_neighbors + 1 and the same seed.
"""
import numpy as np
import pytest

from conftest import latent


def test_fit_transform_layout_on_oracle_engine(monkeypatch):
    import pynndescent_b200.pynndescent_ as mod
    from oracle_engine import OracleEngine
    from pynndescent_b200 import NNDescent, PyNNDescentTransformer
    from sklearn.base import clone
    monkeypatch.setattr(mod, "Engine", OracleEngine)
    X = latent(700, 12, 3)
    tr = PyNNDescentTransformer(n_neighbors=9, random_state=7, n_jobs=8)
    assert clone(tr).get_params() == tr.get_params() and tr.get_params()["search_epsilon"] == 0.1
    Xt = tr.fit_transform(X)
    assert Xt.shape == (700, 700) and Xt.dtype == np.float32 and Xt.format == "csr"
    assert tr.index_.n_neighbors == 10 and tr.n_samples_fit == 700          # sklearn convention: self not counted
    gi, gd = NNDescent(X, n_neighbors=10, random_state=7, n_jobs=8).neighbor_graph
    assert np.array_equal(np.diff(Xt.indptr), np.full(700, 10))
    assert np.array_equal(Xt.indices.reshape(700, 10), gi)                   # rows keep the ascending-distance order
    assert np.array_equal(Xt.data.reshape(700, 10), gd.astype(np.float32))
    assert np.array_equal(Xt.indices.reshape(700, 10)[:, 0], np.arange(700)) and np.all(Xt.data.reshape(700, 10)[:, 0] == 0)


@pytest.mark.gpu
def test_transformer_equivalence_gpu():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from pynndescent_b200 import NNDescent, PyNNDescentTransformer
    X = latent(1200, 16, 9)
    train, test = X[:800], X[800:]
    nnd = NNDescent(train, n_neighbors=16, random_state=42, compressed=False)
    indices, dists = nnd.query(test, k=15, epsilon=0.15)
    order = np.argsort(indices, axis=1)
    indices_sorted = np.take_along_axis(indices, order, axis=1)
    dists_sorted = np.take_along_axis(dists, order, axis=1)
    tr = PyNNDescentTransformer(n_neighbors=15, search_epsilon=0.15, random_state=42).fit(train, compress_index=False)
    Xt = tr.transform(test).sorted_indices()
    assert Xt.shape == (400, 800)
    assert np.all(Xt.indices == indices_sorted.ravel())
    assert np.allclose(Xt.data, dists_sorted.ravel())
    full = tr.fit_transform(train)
    assert full.shape == (800, 800) and np.array_equal(np.diff(full.indptr), np.full(800, 16))
    # compress_index (pynndescent_.py:2160-2172): the construction-time graph goes away, query() keeps answering
    before = nnd.query(test[:20], k=5)
    nnd.compress_index()
    with pytest.raises(AttributeError):
        nnd.neighbor_graph
    after = nnd.query(test[:20], k=5)
    # (the search rng is threaded through successive calls like the reference's, so allow the rare re-drawn tie)
    assert after[0].shape == (20, 5) and np.mean(before[0] == after[0]) >= 0.9
