"""PyNNDescentTransformer -- the sklearn transformer facade of the reference (pynndescent_.py:2556-2863) over the
sm_100a index.  Host-side glue only: fit() builds an NNDescent with n_neighbors + 1 (sklearn does not count a point as
its own neighbour, :2767-2769), transform() turns query() answers -- or the index's own neighbor_graph for
fit_transform() -- into the CSR matrix sklearn's precomputed-neighbours estimators consume.
"""
import time

import numpy as np
from scipy.sparse import coo_array
from sklearn.base import BaseEstimator, TransformerMixin

from .pynndescent_ import NNDescent


def _ts():
    return time.ctime(time.time())


class PyNNDescentTransformer(BaseEstimator, TransformerMixin):
    """Same constructor keywords and defaults as pynndescent.PyNNDescentTransformer (pynndescent_.py:2706-2745)."""

    def __init__(self, n_neighbors=30, metric="euclidean", metric_kwds=None, n_trees=None, leaf_size=None,
                 search_epsilon=0.1, pruning_degree_multiplier=1.5, diversify_prob=1.0, n_search_trees=1, tree_init=True,
                 random_state=None, n_jobs=None, low_memory=True, max_candidates=None, n_iters=None,
                 early_termination_value=0.001, parallel_batch_queries=False, verbose=False):
        # sklearn's estimator contract: every constructor argument is stored verbatim under its own name (get_params / clone)
        given = dict(locals())
        given.pop("self")
        for name, value in given.items():
            setattr(self, name, value)

    # transformer keyword -> NNDescent keyword, where the two differ (pynndescent_.py:2772-2797)
    _RENAMED = {"early_termination_value": "delta"}
    _NOT_FORWARDED = ("n_neighbors", "search_epsilon", "metric_kwds")

    def fit(self, X, compress_index=True):
        """pynndescent_.py:2747-2799.  The index is built with one neighbour more than asked for: sklearn does not count a point
        as its own neighbour (:2767-2769).  `compress_index` is accepted for signature parity; the device index has no compressed
        form (nothing is dropped, query() keeps working either way)."""
        self.n_samples_fit = X.shape[0]
        if self.verbose:
            print(_ts(), "Creating index")
        kw = {self._RENAMED.get(name, name): value for name, value in self.get_params(deep=False).items()
              if name not in self._NOT_FORWARDED}
        self.index_ = NNDescent(X, n_neighbors=self.n_neighbors + 1, metric_kwds=dict(self.metric_kwds or {}),
                                compressed=compress_index, **kw)
        return self

    def _as_neighbor_matrix(self, indices, distances, n_rows):
        """(n_rows, n_samples_fit) CSR with Xt[i, indices[i, j]] = distances[i, j].  Like the reference (:2822-2833) the triplets
        are ATTACHED to an empty float32 COO container: scipy still believes the (empty) container to be in canonical form, so
        tocsr() neither sorts the columns nor merges entries and every row keeps its ascending-distance order -- building the COO
        from (data, (row, col)) would sort each row by column instead."""
        m = coo_array((n_rows, self.n_samples_fit), dtype=np.float32)
        m.data, m.col = distances.ravel(), indices.ravel()
        m.row = np.repeat(np.arange(indices.shape[0], dtype=np.int32), indices.shape[1])
        return m.tocsr()

    def transform(self, X, y=None):
        """pynndescent_.py:2801-2833: `X=None` returns the fitted points' own graph, otherwise query() answers."""
        if X is None:
            indices, distances = self.index_.neighbor_graph
            n_rows = self.n_samples_fit
        else:
            indices, distances = self.index_.query(X, k=self.n_neighbors, epsilon=self.search_epsilon)
            n_rows = X.shape[0]
        if self.verbose:
            print(_ts(), "Constructing neighbor matrix")
        return self._as_neighbor_matrix(indices, distances, n_rows)

    def fit_transform(self, X, y=None, **fit_params):
        """pynndescent_.py:2835-2863 (the diagonal is always explicit: every point is its own first neighbour)."""
        return self.fit(X, compress_index=False).transform(X=None)
