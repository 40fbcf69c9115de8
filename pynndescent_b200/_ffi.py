"""ctypes binding of libnnd_b200.so (include/nnd_b200.h).

The CUDA library is the product: if it is missing or does not load, importing this module
raises -- there is NO CPU fallback anywhere in pynndescent_b200.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnnd_b200.so")


class NNDLibraryError(RuntimeError):
    pass


class Problem(C.Structure):
    """nnd_problem (include/nnd_b200.h)"""
    _fields_ = [("X", C.c_void_p), ("N", C.c_int64), ("d", C.c_int32), ("ld", C.c_int32), ("metric", C.c_int32),
                ("aux", C.c_void_p), ("k", C.c_int32), ("heap_idx", C.c_void_p), ("heap_dist", C.c_void_p),
                ("heap_flag", C.c_void_p)]


class SearchIndex(C.Structure):
    """nnd_search_index (include/nnd_b200.h)"""
    _fields_ = [("data", C.c_void_p), ("N", C.c_int64), ("d", C.c_int32), ("ld", C.c_int32), ("metric", C.c_int32),
                ("aux", C.c_void_p), ("indptr", C.c_void_p), ("indices", C.c_void_p), ("hyperplanes", C.c_void_p),
                ("offsets", C.c_void_p), ("children", C.c_void_p), ("tree_indices", C.c_void_p),
                ("min_distance", C.c_float), ("n_neighbors", C.c_int32), ("normalize", C.c_int32)]


if not os.path.exists(LIB_PATH):
    raise NNDLibraryError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(nvcc -gencode arch=compute_100a,code=sm_100a); pynndescent_b200 has no CPU fallback")
try:
    lib = C.CDLL(LIB_PATH)
except OSError as e:  # pragma: no cover
    raise NNDLibraryError(f"cannot load {LIB_PATH}: {e}") from e

_vp, _i64, _i32, _sz = C.c_void_p, C.c_int64, C.c_int32, C.c_size_t
_pi64, _pi32, _psz = C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_size_t)
_PP = C.POINTER(Problem)

# every symbol declared in include/nnd_b200.h (tests/test_abi.py checks the two lists agree)
SIGNATURES = {
    "nnd_last_error": (C.c_char_p, []),
    "nnd_version": (C.c_int, []),
    "nnd_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "nnd_destroy": (C.c_int, [_vp]),
    "nnd_metric_aux_bytes": (C.c_int, [_i32, _i64, _psz]),
    "nnd_metric_prepare": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "nnd_make_heap": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "nnd_forest_workspace_bytes": (C.c_int, [_i64, _i32, _i32, _psz]),
    "nnd_forest_build": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _pi64, _i32, _i32, _i32, _vp, _vp, _pi64, _pi32,
                                   _vp, _sz, _vp]),
    "nnd_descent_workspace_bytes": (C.c_int, [_i64, _i32, _i32, _i32, _i64, _psz]),
    "nnd_leaf_init": (C.c_int, [_vp, _PP, _vp, _vp, _pi64, _i64, _i32, _i64, _vp, _sz, _pi64, _vp]),
    "nnd_init_random": (C.c_int, [_vp, _PP, _pi64, _vp, _sz, _vp]),
    "nnd_cand_workspace_bytes": (C.c_int, [_i64, _i32, _i32, _i32, _psz]),
    "nnd_build_candidates": (C.c_int, [_vp, _PP, _i32, _pi64, _i32, _i32, _vp, _vp, _vp, _sz, _vp]),
    "nnd_process_candidates": (C.c_int, [_vp, _PP, _i32, _vp, _vp, _i64, _vp, _sz, _pi64, _pi64, _vp]),
    "nnd_deheap_sort": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "nnd_descent": (C.c_int, [_vp, _PP, _i32, _pi64, _i32, _i32, C.c_double, _i64, _vp, _sz, _pi32, _pi64, _pi64, _vp]),
    "nnd_metric_set_param": (C.c_int, [_vp, _i32, _vp, C.c_double, _vp]),
    "nnd_set_row_range": (C.c_int, [_vp, _i64, _i64]),
    "nnd_set_thresholds": (C.c_int, [_vp, _vp]),
    "nnd_set_heap_trusted": (C.c_int, [_vp, _i32]),
    "nnd_cand_local_counts": (C.c_int, [_vp, _PP, _i32, _i32, _i32, _i64, _vp, _sz, _pi64, _vp]),
    "nnd_cand_export": (C.c_int, [_vp, _PP, _i32, _pi64, _i32, _i32, _i32, _i64, _pi64, _vp, _vp, _sz, _vp]),
    "nnd_cand_import_select": (C.c_int, [_vp, _PP, _i32, _i32, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "nnd_join_prepare": (C.c_int, [_vp, _PP, _i32, _vp, _vp, _i64, _vp, _sz, _pi32, _vp]),
    "nnd_join_emit_round": (C.c_int, [_vp, _PP, _i32, _i64, _vp, _sz, _i32, _pi64, _vp]),
    "nnd_push_export": (C.c_int, [_vp, _PP, _i32, _i64, _vp, _sz, _i64, _i64, _i32, _vp, _pi64, _vp]),
    "nnd_apply_records": (C.c_int, [_vp, _PP, _i32, _vp, _i64, _i64, _vp, _sz, _pi64, _vp]),
    "nnd_leaf_fused_range": (C.c_int, [_vp, _PP, _vp, _vp, _i64, _i64, _i32, _vp]),
    "nnd_leaf_range": (C.c_int, [_vp, _PP, _vp, _vp, _i64, _i64, _i64, _i64, _i32, _i64, _vp, _sz, _vp]),
    "nnd_heap_pack_rows": (C.c_int, [_vp, _PP, _vp, _i64, _vp, _vp, _vp, _i32, _vp]),
    "nnd_query_workspace_bytes": (C.c_int, [_i64, _i64, _i32, _psz]),
    "nnd_query": (C.c_int, [_vp, C.POINTER(SearchIndex), _vp, _i64, _i32, C.c_double, _pi64, _vp, _vp, _vp, _sz, _pi64, _vp]),
    "nnd_hub_tree_workspace_bytes": (C.c_int, [_i64, _i32, _psz]),
    "nnd_hub_tree_build": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                     _i64, _pi32, _pi32, _vp, _sz, _vp]),
    "nnd_diversify": (C.c_int, [_vp, _PP, _vp, _vp, _vp]),
    "nnd_search_graph_workspace_bytes": (C.c_int, [_i64, _i32, _psz]),
    "nnd_search_graph_build": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _i32, _vp, _vp, _i64, _pi64, C.POINTER(C.c_float), _vp,
                                        _sz, _vp]),
    "nnd_rng_stream": (C.c_int, [_vp, _pi64, _i64, _i64, _i32, _vp, _vp]),
    "nnd_rng_advance_host": (C.c_int, [_vp, _pi64, _i64]),
    "nnd_pair_distances": (C.c_int, [_vp, _PP, _vp, _vp, _i64, _vp, _vp]),
    "nnd_profile_read": (C.c_int, [_vp, C.POINTER(C.c_double), _i32]),
    "nnd_x_leaf_gram_tc": (C.c_int, [_vp, _PP, _vp, _vp, _i64, _i64, _i32, _i32, _vp, _vp]),
    "nnd_launch_count": (C.c_int64, [_vp]),
    "nnd_stage_profile": (C.c_int, [_vp, _i32]),
    "nnd_stage_read": (C.c_int, [_vp, C.POINTER(C.c_double), _i32]),
}
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here == ABI drift
    _fn.restype = _res
    _fn.argtypes = _args

NND_ERRORS = {-1: ValueError, -2: RuntimeError, -3: MemoryError, -4: NotImplementedError, -5: OverflowError}


def check(status):
    """Turn a status code into the Python exception the reference would raise at that seam."""
    if status != 0:
        msg = lib.nnd_last_error().decode("utf-8", "replace")
        raise NND_ERRORS.get(status, RuntimeError)(f"libnnd_b200: {msg} (status {status})")


def i64_array(values):
    arr = (C.c_int64 * len(values))(*[int(v) for v in values])
    return arr
