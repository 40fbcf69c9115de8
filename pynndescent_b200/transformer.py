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
        self.n_neighbors = n_neighbors
        self.metric = metric
        self.metric_kwds = metric_kwds
        self.n_trees = n_trees
        self.leaf_size = leaf_size
        self.search_epsilon = search_epsilon
        self.pruning_degree_multiplier = pruning_degree_multiplier
        self.diversify_prob = diversify_prob
        self.n_search_trees = n_search_trees
        self.tree_init = tree_init
        self.random_state = random_state
        self.low_memory = low_memory
        self.max_candidates = max_candidates
        self.n_iters = n_iters
        self.early_termination_value = early_termination_value
        self.n_jobs = n_jobs
        self.parallel_batch_queries = parallel_batch_queries
        self.verbose = verbose

    def fit(self, X, compress_index=True):
        """pynndescent_.py:2747-2799.  `compress_index` is accepted for signature parity; the device index has no
        compressed form (nothing is dropped, query() keeps working either way)."""
        self.n_samples_fit = X.shape[0]
        metric_kwds = {} if self.metric_kwds is None else self.metric_kwds
        if self.verbose:
            print(_ts(), "Creating index")
        effective_n_neighbors = self.n_neighbors + 1
        self.index_ = NNDescent(
            X, metric=self.metric, metric_kwds=metric_kwds, n_neighbors=effective_n_neighbors, n_trees=self.n_trees,
            leaf_size=self.leaf_size, pruning_degree_multiplier=self.pruning_degree_multiplier,
            diversify_prob=self.diversify_prob, n_search_trees=self.n_search_trees, tree_init=self.tree_init,
            random_state=self.random_state, low_memory=self.low_memory, max_candidates=self.max_candidates,
            n_iters=self.n_iters, delta=self.early_termination_value, n_jobs=self.n_jobs, compressed=compress_index,
            parallel_batch_queries=self.parallel_batch_queries, verbose=self.verbose)
        return self

    def transform(self, X, y=None):
        """pynndescent_.py:2801-2833: CSR (n_samples_transform, n_samples_fit), Xt[i, j] = distance of neighbour j."""
        if X is None:
            n_samples_transform = self.n_samples_fit
            indices, distances = self.index_.neighbor_graph
        else:
            n_samples_transform = X.shape[0]
            indices, distances = self.index_.query(X, k=self.n_neighbors, epsilon=self.search_epsilon)
        if self.verbose:
            print(_ts(), "Constructing neighbor matrix")
        result = coo_array((n_samples_transform, self.n_samples_fit), dtype=np.float32)
        result.row = np.repeat(np.arange(indices.shape[0], dtype=np.int32), indices.shape[1])
        result.col = indices.ravel()
        result.data = distances.ravel()
        return result.tocsr()

    def fit_transform(self, X, y=None, **fit_params):
        """pynndescent_.py:2835-2863 (the diagonal is always explicit: every point is its own first neighbour)."""
        self.fit(X, compress_index=False)
        return self.transform(X=None)
