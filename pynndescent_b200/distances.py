"""Named-metric plugin surface (mirror of /root/reference/pynndescent/distances.py:2103-2188).

Same string keys as the reference's `named_distances`; each supported key maps to the C enum of
the kernel that computes its "fast alternative" plus the correction applied at the API edge
(`fast_distance_alternatives`, distances.py:2170-2188).  Names the reference knows but the
sm_100a path does not implement raise NotImplementedError -- never a CPU fallback; unknown names
raise the reference's ValueError (pynndescent_.py:1292).
"""
import numpy as np

NND_METRIC_SQEUCLIDEAN, NND_METRIC_ALT_COSINE, NND_METRIC_ALT_DOT, NND_METRIC_CORRELATION = 0, 1, 2, 3
# SURVEY 8f-2 (include/nnd_b200.h)
(NND_METRIC_MANHATTAN, NND_METRIC_CHEBYSHEV, NND_METRIC_CANBERRA, NND_METRIC_HAMMING, NND_METRIC_ALT_JACCARD,
 NND_METRIC_DICE, NND_METRIC_MATCHING, NND_METRIC_KULSINSKI, NND_METRIC_ROGERSTANIMOTO, NND_METRIC_RUSSELLRAO,
 NND_METRIC_SOKALSNEATH, NND_METRIC_YULE, NND_METRIC_ALT_HELLINGER, NND_METRIC_ALT_INNER, NND_METRIC_MINKOWSKI) = range(4, 19)
NND_CORRECT_NONE, NND_CORRECT_SQRT = 0, 1


def correct_alternative_cosine(d):
    """distances.py:704-711: 1 - 2^-d (float64 result like the reference's numba.vectorize)."""
    return 1.0 - np.power(2.0, -np.asarray(d).astype(np.float64))


def correct_alternative_hellinger(d):
    """distances.py:1419-1426: sqrt(1 - 2^-d)."""
    return np.sqrt(1.0 - np.power(2.0, -np.asarray(d).astype(np.float64)))


def true_angular_from_alt_cosine(d):
    """distances.py:1273-1280: 1 - arccos(2^-d) / pi."""
    return 1.0 - np.arccos(np.power(2.0, -np.asarray(d).astype(np.float64))) / np.pi


def correct_alternative_inner_product(d):
    """distances.py:841-853: -1 / d, and 0 for the FLOAT32_MAX sentinel (non-positive inner product)."""
    d = np.asarray(d)
    d64 = d.astype(np.float64)
    return np.where(d >= np.finfo(np.float32).max, 0.0, -1.0 / np.where(d64 == 0.0, 1.0, d64))


# correct_alternative_jaccard (distances.py:329-336) is the same map as correct_alternative_cosine
correct_alternative_jaccard = correct_alternative_cosine

# name -> (kernel metric id, device-side correction, host-side correction or None, angular trees)
gpu_distances = {
    "euclidean": (NND_METRIC_SQEUCLIDEAN, NND_CORRECT_SQRT, None, False),
    "l2": (NND_METRIC_SQEUCLIDEAN, NND_CORRECT_SQRT, None, False),
    "sqeuclidean": (NND_METRIC_SQEUCLIDEAN, NND_CORRECT_NONE, None, False),
    "cosine": (NND_METRIC_ALT_COSINE, NND_CORRECT_NONE, correct_alternative_cosine, True),
    "dot": (NND_METRIC_ALT_DOT, NND_CORRECT_NONE, correct_alternative_cosine, True),
    "correlation": (NND_METRIC_CORRELATION, NND_CORRECT_NONE, None, True),
    # SURVEY 8f-2; angular-tree membership from pynndescent_.py:1074-1085
    "manhattan": (NND_METRIC_MANHATTAN, NND_CORRECT_NONE, None, False),
    "taxicab": (NND_METRIC_MANHATTAN, NND_CORRECT_NONE, None, False),
    "l1": (NND_METRIC_MANHATTAN, NND_CORRECT_NONE, None, False),
    "chebyshev": (NND_METRIC_CHEBYSHEV, NND_CORRECT_NONE, None, False),
    "linfinity": (NND_METRIC_CHEBYSHEV, NND_CORRECT_NONE, None, False),
    "linfty": (NND_METRIC_CHEBYSHEV, NND_CORRECT_NONE, None, False),
    "linf": (NND_METRIC_CHEBYSHEV, NND_CORRECT_NONE, None, False),
    "canberra": (NND_METRIC_CANBERRA, NND_CORRECT_NONE, None, False),
    "hamming": (NND_METRIC_HAMMING, NND_CORRECT_NONE, None, True),
    "jaccard": (NND_METRIC_ALT_JACCARD, NND_CORRECT_NONE, correct_alternative_jaccard, True),
    "dice": (NND_METRIC_DICE, NND_CORRECT_NONE, None, True),
    "matching": (NND_METRIC_MATCHING, NND_CORRECT_NONE, None, False),
    "kulsinski": (NND_METRIC_KULSINSKI, NND_CORRECT_NONE, None, False),
    "rogerstanimoto": (NND_METRIC_ROGERSTANIMOTO, NND_CORRECT_NONE, None, False),
    "sokalmichener": (NND_METRIC_ROGERSTANIMOTO, NND_CORRECT_NONE, None, False),
    "russellrao": (NND_METRIC_RUSSELLRAO, NND_CORRECT_NONE, None, False),
    "sokalsneath": (NND_METRIC_SOKALSNEATH, NND_CORRECT_NONE, None, False),
    "yule": (NND_METRIC_YULE, NND_CORRECT_NONE, None, False),
    "hellinger": (NND_METRIC_ALT_HELLINGER, NND_CORRECT_NONE, correct_alternative_hellinger, True),
    # fast_distance_alternatives (distances.py:2176-2184); neither name is in the angular-tree list (pynndescent_.py:1074-1085)
    # and only "dot" normalises the data (:1101-1103)
    "inner_product": (NND_METRIC_ALT_INNER, NND_CORRECT_NONE, correct_alternative_inner_product, False),
    "true_angular": (NND_METRIC_ALT_COSINE, NND_CORRECT_NONE, true_angular_from_alt_cosine, False),
    # distances.py:138-153; p comes from metric_kwds (metric_param below), default 2
    "minkowski": (NND_METRIC_MINKOWSKI, NND_CORRECT_NONE, None, False),
}

# metrics whose distance function takes one scalar parameter: name -> (keyword, default) of the reference's signature
parametrised = {"minkowski": ("p", 2.0)}


def metric_param(metric, metric_kwds):
    """pynndescent_.py:1064-1065, 1294-1302: the reference passes tuple(metric_kwds.values()) positionally to the distance function.
    Returns the scalar the kernels need (None for the metrics without one); keywords a GPU metric cannot take raise -- never a
    silent drop."""
    vals = list((metric_kwds or {}).values())
    if isinstance(metric, str) and metric in parametrised:
        if len(vals) > 1:
            raise NotImplementedError(f"metric {metric!r} takes one parameter on the sm_100a path, got {metric_kwds!r}")
        v = float(vals[0]) if vals else parametrised[metric][1]
        if not (v > 0.0 and v < 1.0e6):
            raise ValueError(f"metric {metric!r}: parameter {v!r} is outside (0, 1e6)")
        return v
    if vals:
        raise NotImplementedError(f"metric_kwds {metric_kwds!r} are not supported for metric {metric!r} on the sm_100a path")
    return None

# every key of the reference's named_distances (distances.py:2103-2162)
named_distances = (
    "euclidean", "l2", "sqeuclidean", "manhattan", "taxicab", "l1", "chebyshev", "linfinity", "linfty", "linf",
    "minkowski", "seuclidean", "standardised_euclidean", "wminkowski", "weighted_minkowski", "mahalanobis",
    "canberra", "cosine", "dot", "inner_product", "correlation", "haversine", "braycurtis", "spearmanr", "tsss",
    "true_angular", "hellinger", "kantorovich", "wasserstein", "wasserstein_1d", "wasserstein-1d", "kantorovich-1d",
    "kantorovich_1d", "circular_kantorovich", "circular_wasserstein", "sinkhorn", "jensen-shannon", "jensen_shannon",
    "symmetric-kl", "symmetric_kl", "symmetric_kullback_liebler", "hamming", "jaccard", "dice", "matching",
    "kulsinski", "rogerstanimoto", "russellrao", "sokalsneath", "sokalmichener", "yule", "bit_hamming", "bit_jaccard",
)


def resolve(metric):
    """pynndescent_.py:1271-1292 _set_distance_func resolution order, GPU edition."""
    if callable(metric):
        raise NotImplementedError("callable (numba) metrics cannot run on the sm_100a path; "
                                  "use one of: " + ", ".join(sorted(gpu_distances)))
    if metric in gpu_distances:
        return gpu_distances[metric]
    if metric in named_distances or (isinstance(metric, str) and metric.startswith("proxy_")):
        raise NotImplementedError(f"metric {metric!r} is not implemented on the sm_100a path (no CPU fallback); "
                                  "supported: " + ", ".join(sorted(gpu_distances)))
    raise ValueError("Metric is neither callable, nor a recognised string")
