"""CPU-only: the C-ABI library loads and exports exactly what include/nnd_b200.h declares."""
import os
import re
import subprocess

import numpy as np

from conftest import ROOT


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "nnd_b200.h")).read()
    return sorted(set(re.findall(r"^NND_API [\w \*]+?\b(nnd_\w+)\(", src, flags=re.M)))


def test_header_declares_symbols():
    syms = _header_symbols()
    assert len(syms) >= 20 and "nnd_descent" in syms and "nnd_forest_build" in syms


def test_library_exports_every_declared_symbol():
    so = os.path.join(ROOT, "pynndescent_b200", "libnnd_b200.so")
    assert os.path.exists(so), "libnnd_b200.so missing: run __graft_entry__.build()"
    out = subprocess.check_output(["nm", "-D", "--defined-only", so], text=True)
    exported = set(re.findall(r" T (nnd_\w+)", out))
    assert set(_header_symbols()) <= exported, sorted(set(_header_symbols()) - exported)
    # nothing but the declared C-ABI leaks out
    assert exported <= set(_header_symbols()), sorted(exported - set(_header_symbols()))


def test_ctypes_binding_covers_header():
    from pynndescent_b200 import _ffi
    assert sorted(_ffi.SIGNATURES) == _header_symbols()
    assert _ffi.lib.nnd_version() >= 100


def test_metric_registry_matches_reference_names():
    from pynndescent_b200 import distances as dist
    import pytest
    assert dist.resolve("euclidean")[0] == dist.NND_METRIC_SQEUCLIDEAN
    assert dist.resolve("cosine")[3] is True
    with pytest.raises(ValueError, match="neither callable, nor a recognised string"):
        dist.resolve("no-such-metric")
    # SURVEY 8f-2 metrics: ids, aliases, angular-tree membership (pynndescent_.py:1074-1085), corrections
    assert dist.resolve("manhattan")[0] == dist.resolve("l1")[0] == dist.resolve("taxicab")[0] == dist.NND_METRIC_MANHATTAN
    assert dist.resolve("linf")[0] == dist.NND_METRIC_CHEBYSHEV and dist.resolve("sokalmichener")[0] == dist.NND_METRIC_ROGERSTANIMOTO
    assert [dist.resolve(m)[3] for m in ("jaccard", "dice", "hellinger", "hamming", "manhattan", "yule")] == [True] * 4 + [False] * 2
    assert dist.resolve("jaccard")[2] is not None and dist.resolve("hellinger")[2] is not None and dist.resolve("dice")[2] is None
    # fast_distance_alternatives (distances.py:2176-2184): inner_product has its own tail in the dot kernel class, true_angular
    # is the alternative_cosine kernel under another correction; neither uses angular trees
    assert dist.resolve("inner_product")[0] == dist.NND_METRIC_ALT_INNER == 17
    assert dist.resolve("true_angular")[0] == dist.NND_METRIC_ALT_COSINE
    assert dist.resolve("inner_product")[3] is False and dist.resolve("true_angular")[3] is False
    assert dist.resolve("inner_product")[2](np.array([0.5, 3.4028235e38], np.float32)).tolist() == [-2.0, 0.0]
    assert dist.resolve("true_angular")[2](np.array([0.0, 1.0], np.float32)).tolist() == [1.0, 1.0 - np.arccos(0.5) / np.pi]
    assert dist.resolve("minkowski")[0] == dist.NND_METRIC_MINKOWSKI == 18 and dist.metric_param("minkowski", {"p": 3}) == 3.0
    with pytest.raises(NotImplementedError):
        dist.resolve("wminkowski")         # needs a weight vector
    with pytest.raises(NotImplementedError):
        dist.resolve("bit_jaccard")
    with pytest.raises(NotImplementedError):
        dist.resolve(lambda a, b: 0.0)


def test_host_only_entry_points_without_a_gpu():
    """The sizing / registry entry points are host code: they answer without a device (no compute call is made here)."""
    import ctypes as C
    from pynndescent_b200 import _ffi, distances as dist
    nb = C.c_size_t()
    for mid, per_row in ((dist.NND_METRIC_SQEUCLIDEAN, 0), (dist.NND_METRIC_ALT_COSINE, 4), (dist.NND_METRIC_CORRELATION, 16),
                         (dist.NND_METRIC_ALT_INNER, 0), (dist.NND_METRIC_MINKOWSKI, 0)):
        assert _ffi.lib.nnd_metric_aux_bytes(mid, 1000, C.byref(nb)) == 0
        assert nb.value == 1000 * per_row + 16          # the 16 bytes every aux buffer has hold minkowski's (p, 1/p)
    assert _ffi.lib.nnd_metric_aux_bytes(19, 1000, C.byref(nb)) != 0
    assert b"unknown metric" in _ffi.lib.nnd_last_error()
    fb = C.c_size_t()
    assert _ffi.lib.nnd_forest_workspace_bytes(100000, 128, 12, C.byref(fb)) == 0
    assert fb.value > 12 * 100000 * 4 * 8               # perm, seg, act (x2 each), flags, scan, invp ... per (tree, point)


def test_sm100a_sass_has_bulk_copy():
    """The join kernel stages rows with the TMA engine's bulk copy: SASS must show UBLKCP."""
    so = os.path.join(ROOT, "pynndescent_b200", "libnnd_b200.so")
    try:
        sass = subprocess.check_output(["cuobjdump", "-sass", so], text=True, stderr=subprocess.DEVNULL)
    except (OSError, subprocess.CalledProcessError):
        import pytest
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in sass or "SM100" in sass.upper()
    assert "UBLKCP" in sass
