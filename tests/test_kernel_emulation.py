"""CPU emulation of the SOURCE of the CUDA kernels (tests/emu/cuda_emu.h).

The kernels are cut out of pynndescent_b200/csrc/*.cu by their section markers (the test fails loudly if a marker disappears),
compiled with g++ against a small emulation of the CUDA execution model (one OS thread per CUDA thread, real barriers and
shuffles, PTX-semantics mbarriers, an asynchronous "TMA" thread serving the bulk copies) and driven through the same host
sequences as the library.  Checked against the oracle bit for bit: the batched local join half-update by half-update in emission
order, whole iterations (join -> counting sort -> ordered replay with its filters), the candidate kernels, leaf init (fused and
split), the forest (position- and point-major passes), the query walk, init_random / deheap_sort / diversification, the
search-graph kernels and the hub tree.

It started in round 1 as a stand-in for GPU time; every kernel it covers has since run on hardware (`-m gpu` suite), and it stays
as the CPU-side regression net for the kernel sources.  This is an emulation of the programming model, not of the hardware: it
proves indexing, tiling, protocols and the arithmetic order, not performance and not memory-model subtleties of the async proxy.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, latent

# a deadlocked emulation must fail, not hang the suite (mbarrier waits abort after 4 minutes on their own; plain barriers cannot)
pytestmark = pytest.mark.timeout(1500, method="thread")

CSRC = os.path.join(ROOT, "pynndescent_b200", "csrc")
EMU = os.path.join(ROOT, "tests", "emu")


def _between(text, start, end, what):
    a = text.find(start)
    b = text.find(end, a + 1)
    assert a >= 0 and b > a, f"section markers of {what} not found -- update tests/test_kernel_emulation.py"
    return text[a:b]


@pytest.fixture(scope="module")
def emu_lib(tmp_path_factory):
    tmp = str(tmp_path_factory.mktemp("emu"))
    common = open(os.path.join(CSRC, "nnd_common.cuh")).read()
    ptx = _between(common, "// ------------------------------------------------------------------ PTX: mbarrier",
                   "// ------------------------------------------------------------------ host-visible kernels", "the PTX helpers")
    common = common.replace(ptx, "")
    for inc in ("#include <cuda_runtime.h>\n", "#include <math_constants.h>\n"):
        assert inc in common
        common = common.replace(inc, "")
    common = common.replace('#include "../../include/nnd_b200.h"', '#include "%s"' % os.path.join(ROOT, "include", "nnd_b200.h"))
    open(os.path.join(tmp, "nnd_common.cuh"), "w").write(common)
    open(os.path.join(tmp, "join_batch.cuh"), "w").write(open(os.path.join(CSRC, "join_batch.cuh")).read())
    desc = open(os.path.join(CSRC, "k_descent.cu")).read()
    batch = _between(desc, "// prepass 1: per row -> active flag", "// ------------------------------------------------------------------ routing",
                     "k_jb_* / k_join_batch")
    cand = open(os.path.join(CSRC, "k_cand.cu")).read()
    cand_sh = _between(cand, "struct __align__(4) CandRec", "static int cand_range(", "the partitioned / row-sharded candidate kernels")
    cand = _between(cand, "__global__ void k_cand_count", "__global__ void k_rng_table", "the candidate kernels")
    sg = open(os.path.join(CSRC, "k_search_graph.cu")).read()
    sg = _between(sg, "#define SG_WARPS", "// ------------------------------------------------------------------------------------------------ host",
                  "the search-graph kernels")
    route = _between(desc, "// ------------------------------------------------------------------ routing",
                     "__global__ void k_row_bound(", "k_scatter / k_scatter_rank / k_apply")
    join = _between(desc, "#define DK 128", "// =====================================================================================================\n// k_join_batch:",
                    "JoinParams / k_join")
    leaf = _between(desc, "// ------------------------------------------------------------------ split leaf init",
                    "// ------------------------------------------------------------------ host orchestration", "k_leaf_prep / k_leaf_push")
    forest = open(os.path.join(CSRC, "k_forest.cu")).read()
    forest = _between(forest, "// ---------------------------------------------------------------- canonical warp reductions",
                      'extern "C" int nnd_forest_build', "the forest kernels")
    forest = forest.replace("extern __shared__ float s_t[];", "float *s_t = (float *)emu::dyn_smem();")
    query = open(os.path.join(CSRC, "k_query.cu")).read()
    query = _between(query, "struct Seed {", 'extern "C" int nnd_query_workspace_bytes', "the query kernels")
    misc = open(os.path.join(CSRC, "k_misc.cu")).read()
    misc = (_between(misc, "// ------------------------------------------------------------------ init_random", "int nnd_rng_table_launch", "k_init_random") +
            _between(misc, "// ------------------------------------------------------------------ deheap_sort", 'extern "C" int nnd_deheap_sort', "k_deheap_sort"))
    prep = open(os.path.join(CSRC, "k_prepare.cu")).read()
    misc += _between(prep, "// ------------------------------------------------------------------ diversification", 'extern "C" int nnd_diversify', "k_diversify")
    hub = _between(prep, "// ------------------------------------------------------------------ canonical warp reductions (as in k_forest.cu)",
                   'extern "C" int nnd_hub_tree_workspace_bytes', "the hub-tree kernels")
    hub = hub.replace("extern __shared__ float s_t[];", "float *s_t = (float *)emu::dyn_smem();")
    kernels = (join + "\n" + batch + "\n#define EMU_WITH_CAND 1\n" + cand + "\n#define EMU_WITH_CANDSH 1\n" + cand_sh + "\n#define EMU_WITH_SG 1\n" + sg +
               "\n#define EMU_WITH_APPLY 1\n" + route + "\n#define EMU_WITH_LEAF 1\n" + leaf + "\n#define EMU_WITH_FOREST 1\n" + forest +
               "\n#define EMU_WITH_QUERY 1\n" + query + "\n#define EMU_WITH_MISC 1\n" + misc +
               "\n#define EMU_WITH_HUB 1\n" + hub)
    n_ext = kernels.count("extern __shared__")
    for al in ("16", "128"):
        kernels = kernels.replace("extern __shared__ __align__(%s) unsigned char smem_raw[];" % al, "unsigned char *smem_raw = emu::dyn_smem();")
    assert n_ext == 11 and "extern __shared__" not in kernels
    src = os.path.join(tmp, "emu_kernels.cpp")
    with open(src, "w") as f:
        f.write('#include "%s"\n#include "join_batch.cuh"\n' % os.path.join(EMU, "cuda_emu.h"))
        f.write(kernels)
        f.write(open(os.path.join(EMU, "emu_driver.inc")).read())
    so = os.path.join(tmp, "libemu.so")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-g", "-pthread", "-ffp-contract=off", "-fno-fast-math", "-mfma", "-shared", "-fPIC",
                           "-I", tmp, "-o", so, src])
    lib = C.CDLL(so)
    f32 = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS"); i32 = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
    u32 = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS"); u64 = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
    i64 = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
    lib.emu_join.argtypes = [f32, C.c_long, C.c_int, C.c_int, C.c_void_p, f32, i32, i32, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                             C.c_int, C.c_long, i32, i32, f32, u32, u64, u32, i64]
    lib.emu_norms.argtypes = [f32, C.c_long, C.c_int, f32]
    u8 = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
    lib.emu_cand.argtypes = [i32, u8, C.c_long, C.c_int, C.c_int, C.c_int, i64, C.c_int, i32, i32, C.c_long, C.c_long, C.c_int]
    lib.emu_cand_sharded.argtypes = [i32, u8, C.c_long, C.c_int, C.c_int, C.c_int, i64, C.c_int, C.c_int, i32, i32]
    lib.emu_search_graph.argtypes = [i32, f32, i32, f32, C.c_long, C.c_int, i32, C.c_int, i32, i32, f32]
    lib.emu_search_graph.restype = C.c_long
    lib.emu_process_candidates.argtypes = [f32, C.c_long, C.c_int, C.c_int, C.c_void_p, i32, f32, u8, C.c_int, i32, i32, C.c_int, C.c_void_p,
                                           C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.emu_process_candidates.restype = C.c_long
    lib.emu_leaf_init.argtypes = [f32, C.c_long, C.c_int, C.c_int, i32, i64, C.c_long, C.c_int, i32, f32, u8, C.c_int, C.c_int]
    lib.emu_leaf_init_metric.argtypes = [f32, C.c_long, C.c_int, C.c_int, C.c_int, C.c_void_p, i32, i64, C.c_long, C.c_int, i32, f32, u8, C.c_int]
    lib.emu_pair_distances.argtypes = [f32, C.c_int, C.c_int, C.c_int, C.c_void_p, i32, i32, C.c_long, f32]
    lib.emu_pair_distances.restype = None
    lib.emu_query.argtypes = [f32, C.c_long, C.c_int, C.c_int, C.c_int, C.c_void_p, i32, i32, f32, f32, i32, i32, f32, C.c_long, C.c_int, C.c_double,
                              C.c_float, C.c_int, C.c_int, i64, i32, f32, i64]
    lib.emu_init_random.argtypes = [f32, C.c_long, C.c_int, C.c_int, C.c_int, C.c_void_p, i32, f32, u8, C.c_int, i64]
    lib.emu_init_random.restype = C.c_long
    lib.emu_deheap_sort.argtypes = [i32, f32, C.c_long, C.c_int, C.c_int, f32]
    lib.emu_deheap_sort.restype = None
    lib.emu_diversify.argtypes = [f32, C.c_long, C.c_int, C.c_int, C.c_int, C.c_void_p, i32, f32, C.c_int]
    lib.emu_diversify.restype = None
    lib.emu_hub_tree.argtypes = [f32, C.c_long, C.c_int, C.c_int, i32, C.c_int, C.c_int, C.c_int, i32, i32, i32, i32, i32, i32, f32, f32, C.c_long, i32]
    lib.emu_forest.argtypes = [f32, C.c_long, C.c_int, C.c_int, i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_long, i32, i64, i64]
    return lib


class Case:
    """One local-join input (heap thresholds + candidate lists from the oracle) and the expected half-updates."""

    def __init__(self, n, d, k, metric, seed):
        from sklearn.utils import check_random_state
        from oracle import nnd_oracle as orc
        self.metric, self.C = metric, k
        X = latent(n, d, seed)
        self.ld = (d + 3) // 4 * 4
        self.X = np.zeros((n, self.ld), np.float32); self.X[:, :d] = X
        rs = check_random_state(seed)
        rng_state = rs.randint(orc.INT32_MIN, orc.INT32_MAX, 3).astype(np.int64)
        la, _ = orc.make_forest(X, k, 3, None, rs, metric == "cosine", 200, "level")
        heap = orc.make_heap(n, k)
        orc.init_rp_tree(X, metric, heap, la); orc.init_random(X, metric, heap, rng_state)
        newc, oldc = orc.new_build_candidates(heap, k, rng_state, 8)
        orc.process_candidates(X, metric, heap, newc, oldc, 8)             # second iteration: mixed new / old lists, holes
        self.newc, self.oldc = orc.new_build_candidates(heap, k, rng_state, 8)
        self.thr = np.ascontiguousarray(heap[1][:, 0], np.float32)
        # expected half-updates in emission order (utils.py:575-658 loop order, the kernels' sequence key)
        exp = []
        stride = 2 * k
        for i in range(n):
            nw = self.newc[i][self.newc[i] >= 0]; ol = self.oldc[i][self.oldc[i] >= 0]
            mem = np.concatenate([nw, ol]); nn = len(nw)
            for a in range(nn):
                for b in range(a, len(mem)):
                    p, q = int(mem[a]), int(mem[b])
                    dd = np.float32(orc.dist(metric, X[p], X[q]))
                    base = (i << 20) | ((a * stride + b) << 1)
                    if dd < self.thr[p]:
                        exp.append((base, p, q, dd))
                    if b != a and dd < self.thr[q]:
                        exp.append((base | 1, q, p, dd))
        self.expected = exp
        self.n = n


@pytest.fixture(scope="module")
def cases():
    return {"euclid40": Case(500, 40, 10, "euclidean", 3), "euclid128": Case(300, 128, 15, "euclidean", 4),
            "cosine72": Case(400, 72, 10, "cosine", 5),
            # C = 30: one row can outweigh a whole batch window, so some windows are EMPTY (descriptor with 0 units)
            "euclid24_k30": Case(160, 24, 30, "euclidean", 6),
            "euclid16_tiny": Case(60, 16, 4, "euclidean", 7),        # one d-chunk, 4 candidates, a handful of batches
            "cosine8": Case(120, 7, 6, "cosine", 8),                 # padded rows (d = 7 -> ld = 8), 32-byte bulk copies
            # the dot class with both tails: -log2(dot) and 1/dot (inner_product); mixed-sign rows, so about half of all pairs
            # carry the FLOAT32_MAX sentinel (never below a finite threshold, below an infinite one)
            "dot40": Case(400, 40, 10, "dot", 9), "inner40": Case(400, 40, 10, "inner_product", 10)}


def _run(lib, case, kernel, dk=32, nst=4, grid=3, rank_scatter=0, order=None, capacity=None):
    n = case.n
    cap = capacity or (len(case.expected) + 64)
    aux = None
    if case.metric == "cosine":
        aux = np.zeros(n, np.float32); lib.emu_norms(case.X, n, case.ld, aux)
    tgt = np.zeros(cap, np.int32); oth = np.zeros(cap, np.int32); dd = np.zeros(cap, np.float32)
    pos = np.full(cap, 0xFFFFFFFF, np.uint32); seq = np.zeros(cap, np.uint64); cnt = np.zeros(n + 1, np.uint32)
    info = np.zeros(8, np.int64)
    rc = lib.emu_join(case.X, n, case.ld, {"euclidean": 0, "cosine": 1, "dot": 2, "inner_product": 17}[case.metric], None if aux is None else aux.ctypes.data_as(C.c_void_p),
                      case.thr, case.newc, case.oldc, case.C, None if order is None else order.ctypes.data_as(C.c_void_p), kernel, dk, nst,
                      grid, rank_scatter, cap, tgt, oth, dd, pos, seq, cnt, info)
    assert rc == 0
    return dict(tgt=tgt, oth=oth, d=dd, pos=pos, seq=seq, cnt=cnt, n_active=int(info[0]), n_batches=int(info[1]), n_out=int(info[2]),
                overflow=int(info[3]), rows=int(info[4]), pairs=int(info[5]))


def _check(case, r, rank_scatter=False):
    exp = case.expected
    assert r["overflow"] == 0 and r["n_out"] == len(exp)
    m = r["n_out"]
    order = np.argsort(r["seq"][:m], kind="stable")
    got = [(int(r["seq"][j]), int(r["tgt"][j]), int(r["oth"][j]), np.float32(r["d"][j])) for j in order]
    assert len({g[0] for g in got}) == m                                    # emission keys are unique
    assert got == [(int(s), p, q, dd) for s, p, q, dd in exp]               # same records, same order, same distance bits
    cnt = np.bincount(r["tgt"][:m], minlength=case.n)
    assert np.array_equal(cnt, r["cnt"][: case.n])                          # per-target counts (RED or ATOM path)
    if not rank_scatter:                                                    # ranks inside every target's segment: a permutation
        for t in np.unique(r["tgt"][:m])[:200]:
            assert sorted(r["pos"][:m][r["tgt"][:m] == t].tolist()) == list(range(int(cnt[t])))
    mem = sum(int((case.newc[i] >= 0).sum() + (case.oldc[i] >= 0).sum()) for i in range(case.n) if (case.newc[i] >= 0).any())
    assert r["rows"] == mem                                                 # profile counter: rows gathered


@pytest.mark.parametrize("name", ["euclid40", "euclid128", "cosine72", "euclid24_k30", "dot40", "inner40"])
def test_emulator_reproduces_the_hardware_verified_kernel(emu_lib, cases, name):
    _check(cases[name], _run(emu_lib, cases[name], kernel=0))


def test_rank_scatter_red_path(emu_lib, cases):
    case = cases["euclid40"]
    _check(case, _run(emu_lib, case, kernel=0, rank_scatter=1), rank_scatter=True)


def test_cases_are_not_trivial(emu_lib, cases):
    for name, case in cases.items():
        r = _run(emu_lib, case, kernel=0, grid=2)
        if case.n >= 150:
            assert r["n_batches"] >= 12 and r["n_out"] >= 2000 and r["n_active"] > case.n // 4, (name, r["n_batches"], r["n_out"], r["n_active"])
        print(name, "batches", r["n_batches"], "active rows", r["n_active"], "half-updates", r["n_out"], "pairs", r["pairs"])


@pytest.mark.parametrize("world,skip", [(1, 0), (2, 1), (3, 0), (5, 1)])
def test_partitioned_candidate_kernels_for_simulated_processes(emu_lib, world, skip):
    """The row-sharded candidate build's kernels (k_cand_count_sh / k_cand_bucket_scan / k_cand_locals / k_cand_export /
    k_cand_import_* + the selection kernels) under the emulator, `world` processes simulated on one heap with the host sequence of
    sharded.py (count tables summed and offset, records routed by owner): identical lists and flag clearing to the oracle's
    new_build_candidates over three iterations of a small build."""
    from sklearn.utils import check_random_state
    from oracle import nnd_oracle as orc
    n, d, k = 700, 16, 8
    X = latent(n, d, 12)
    rs = check_random_state(4)
    rng_state = rs.randint(orc.INT32_MIN, orc.INT32_MAX, 3).astype(np.int64)
    la, _ = orc.make_forest(X, k, 3, None, rs, False, 200, "level")
    heap = orc.make_heap(n, k)
    orc.init_rp_tree(X, "euclidean", heap, la); orc.init_random(X, "euclidean", heap, rng_state)
    for it in range(3):
        ind = heap[0].copy(); flags_emu = heap[2].copy()
        en = np.full((n, k), -7, np.int32); eo = np.full((n, k), -7, np.int32)
        assert emu_lib.emu_cand_sharded(ind, flags_emu, n, k, k, 8, rng_state, world, skip, en, eo) == 0
        newc, oldc = orc.new_build_candidates(heap, k, rng_state, 8)
        assert np.array_equal(en, newc)
        has_new = (newc >= 0).any(1)
        assert np.array_equal(eo[has_new], oldc[has_new])
        assert np.all(eo[~has_new] == -1) if skip else np.array_equal(eo, oldc)
        assert np.array_equal(flags_emu, heap[2])
        orc.process_candidates(X, "euclidean", heap, newc, oldc, 8)


@pytest.mark.parametrize("skip,keys", [(0, 0), (1, 0)])
def test_candidate_kernels_and_cand_skip(emu_lib, skip, keys):
    """k_cand_count / k_cand_fill / k_cand_select* under the emulator against the oracle's new_build_candidates, over three
    iterations of a small build (hub vertices with > 32 reverse edges included).  skip=0 is the hardware-verified schedule
    (all rows must match); skip=1 = NND_CAND_SKIP: rows WITH new candidates must match exactly, the others must come out
    empty, and the sampled-flag clearing must be identical."""
    from sklearn.utils import check_random_state
    from oracle import nnd_oracle as orc
    n, d, k = 900, 20, 10
    X = latent(n, d, 8)
    rs = check_random_state(2)
    rng_state = rs.randint(orc.INT32_MIN, orc.INT32_MAX, 3).astype(np.int64)
    la, _ = orc.make_forest(X, k, 3, None, rs, False, 200, "level")
    heap = orc.make_heap(n, k)
    orc.init_rp_tree(X, "euclidean", heap, la); orc.init_random(X, "euclidean", heap, rng_state)
    saw_inactive = saw_hub = False
    for it in range(4):
        ind = heap[0].copy(); flags_emu = heap[2].copy()
        indeg = np.bincount(ind[ind >= 0], minlength=n)
        saw_hub |= bool((indeg > 32).any())
        en = np.empty((n, k), np.int32); eo = np.empty((n, k), np.int32)
        emu_lib.emu_cand(ind, flags_emu, n, k, k, 8, rng_state, skip, en, eo, 0, 2 ** 62, keys)
        newc, oldc = orc.new_build_candidates(heap, k, rng_state, 8)         # also clears the sampled flags of `heap`
        assert np.array_equal(en, newc)
        has_new = (newc >= 0).any(1)
        saw_inactive |= bool((~has_new).any())
        assert np.array_equal(eo[has_new], oldc[has_new])
        if skip:
            assert np.all(eo[~has_new] == -1)
        else:
            assert np.array_equal(eo, oldc)
        assert np.array_equal(flags_emu, heap[2])
        orc.process_candidates(X, "euclidean", heap, newc, oldc, 8)
    assert saw_hub and saw_inactive


def _sg_emulate(emu_lib, di, dd, di2, dd2, vertex_order, max_degree):
    n, k = di.shape
    indptr = np.zeros(n + 1, np.int32); indices = np.zeros(2 * n * k + 8, np.int32); md = np.zeros(1, np.float32)
    nnz = emu_lib.emu_search_graph(np.ascontiguousarray(di, np.int32), np.ascontiguousarray(dd, np.float32), np.ascontiguousarray(di2, np.int32),
                                   np.ascontiguousarray(dd2, np.float32), n, k, np.ascontiguousarray(vertex_order, np.int32), max_degree, indptr, indices, md)
    return indptr, indices[:nnz], md[0]


@pytest.mark.parametrize("name", ["query_cosine_2k24", "query_jaccard_2k40"])
def test_search_graph_kernels_match_reference_structures(emu_lib, name):
    """k_search_graph.cu under the emulator, on the two diversification passes of the reference's own graphs: the CSR must be
    the reference's search graph (query goldens) -- same check as tests/test_prepare_host.py makes for the host restatement."""
    from oracle import nnd_oracle as orc
    from conftest import load_golden
    from test_prepare_host import _diversify_passes
    G = load_golden(name)
    di, dd, di2, dd2 = _diversify_passes(orc, G, packed=True)
    indptr, indices, md = _sg_emulate(emu_lib, di, dd, di2, dd2, G["vertex_order"], int(np.round(1.5 * int(G["n_neighbors"]))))
    assert np.array_equal(indptr, G["indptr"]) and np.array_equal(indices, G["indices"])
    assert np.float32(md) == np.float32(G["min_distance"])


def test_search_graph_kernels_hubs_ties_and_big_rows(emu_lib):
    """Rows beyond SG_CAP entries (sorted in global scratch), weight ties at the degree cut, rows with only reverse edges,
    empty rows, self loops -- against the host restatement."""
    from pynndescent_b200.prepare import search_graph_bookkeeping
    rs = np.random.RandomState(4)
    n, k, nn = 1500, 12, 6
    di = np.full((n, k), -1, np.int32); dd = np.full((n, k), np.inf, np.float32)
    for i in range(n):
        if i % 7 == 3:
            continue                                          # rows without forward edges
        m = rs.randint(1, k + 1)
        nb = rs.choice(n, m, replace=False)
        nb[: (2 * m) // 3] = rs.choice(3, (2 * m) // 3)        # vertices 0..2 collect > 256 reverse edges each
        nb = np.array(list(dict.fromkeys(int(v) for v in nb)), np.int32)
        if rs.random_sample() < 0.1 and i not in nb:
            nb[-1] = i                                        # a self loop (dropped with the diagonal)
        w = np.sort(np.round(rs.random_sample(len(nb)), 1).astype(np.float32) + np.float32(0.1))
        di[i, : len(nb)] = nb; dd[i, : len(nb)] = w
    di2 = di.copy()
    drop = (rs.random_sample(di.shape) < 0.1) & (di >= 0); di2[drop] = -1      # second-pass drops (holes inside rows)
    vo = rs.permutation(n).astype(np.int32)
    S = search_graph_bookkeeping(di, dd, di2, dd, vo, nn)
    indptr, indices, md = _sg_emulate(emu_lib, di, dd, di2, dd, vo, int(np.round(1.5 * nn)))
    rev_deg = np.bincount(di2[di2 >= 0], minlength=n)
    assert rev_deg.max() > 300 and (np.diff(S["indptr"]) == 0).any()
    assert np.array_equal(indptr, S["indptr"]) and np.array_equal(indices, S["indices"]) and np.float32(md) == S["min_distance"]
    # the layout nnd_diversify really hands over: second-pass survivors packed to the left with their own distances
    order = np.argsort(di2 < 0, axis=1, kind="stable")
    pi = np.take_along_axis(di2, order, 1); pd = np.take_along_axis(dd, order, 1); pd[pi < 0] = np.inf
    indptr, indices, md = _sg_emulate(emu_lib, di, dd, pi, pd, vo, int(np.round(1.5 * nn)))
    assert np.array_equal(indptr, S["indptr"]) and np.array_equal(indices, S["indices"]) and np.float32(md) == S["min_distance"]


@pytest.mark.parametrize("skip", [0, 1])
def test_candidate_kernels_row_sharded(emu_lib, skip):
    """The sharded build (nnd_set_row_range): every process samples candidates for its own row range only, from the replicated
    heap.  Two ranges, each emulated separately, must add up to the oracle's lists (and flag clearing) -- with and without
    NND_CAND_SKIP."""
    from sklearn.utils import check_random_state
    from oracle import nnd_oracle as orc
    n, d, k = 700, 16, 8
    X = latent(n, d, 12)
    rs = check_random_state(4)
    rng_state = rs.randint(orc.INT32_MIN, orc.INT32_MAX, 3).astype(np.int64)
    la, _ = orc.make_forest(X, k, 3, None, rs, False, 200, "level")
    heap = orc.make_heap(n, k)
    orc.init_rp_tree(X, "euclidean", heap, la); orc.init_random(X, "euclidean", heap, rng_state)
    for it in range(3):
        ind = heap[0].copy(); f0 = heap[2].copy()
        B = n // 2 + 1                                             # utils.py:259 ownership rule for 2 processes
        parts = []
        for lo, hi in ((0, B), (B, n)):
            fl = f0.copy(); en = np.full((n, k), -7, np.int32); eo = np.full((n, k), -7, np.int32)
            emu_lib.emu_cand(ind, fl, n, k, k, 8, rng_state, skip, en, eo, lo, hi, 2 * skip)
            assert np.all(en[:lo] == -7) and np.all(en[hi:] == -7)   # nothing written outside the own range
            parts.append((lo, hi, en, eo, fl))
        newc, oldc = orc.new_build_candidates(heap, k, rng_state, 8)
        has_new = (newc >= 0).any(1)
        for lo, hi, en, eo, fl in parts:
            assert np.array_equal(en[lo:hi], newc[lo:hi])
            sel = has_new[lo:hi]
            assert np.array_equal(eo[lo:hi][sel], oldc[lo:hi][sel])
            if not skip:
                assert np.array_equal(eo[lo:hi], oldc[lo:hi])
            assert np.array_equal(fl[lo:hi], heap[2][lo:hi]) and np.array_equal(fl[:lo], f0[:lo]) and np.array_equal(fl[hi:], f0[hi:])
        orc.process_candidates(X, "euclidean", heap, newc, oldc, 8)


@pytest.mark.parametrize("kernel,rank_scatter,leaf_order,short_sort", [(0, 0, False, 0), (0, 1, False, 1)])
def test_whole_iterations_match_the_oracle_heaps(emu_lib, kernel, rank_scatter, leaf_order, short_sort):
    """Three NN-descent iterations: candidates from the oracle, then join -> scatter -> apply under the emulator; after every
    iteration the heaps (indices, distances, flags, layout) and the accepted-push count must equal the oracle's
    process_candidates.  (0, 0) is the hardware-verified schedule = emulator check; the others are the switches of DESIGN 10a
    (NND_RANK_SCATTER; short_sort bit 0 = NND_APPLY_SHORT_SORT; the replay filters of k_apply are always on)."""
    from sklearn.utils import check_random_state
    from oracle import nnd_oracle as orc
    n, d, k = 450, 24, 10
    X = latent(n, d, 21)
    rs = check_random_state(6)
    rng_state = rs.randint(orc.INT32_MIN, orc.INT32_MAX, 3).astype(np.int64)
    la, _ = orc.make_forest(X, k, 3, None, rs, False, 200, "level")
    heap = orc.make_heap(n, k)
    orc.init_rp_tree(X, "euclidean", heap, la); orc.init_random(X, "euclidean", heap, rng_state)
    order = np.random.RandomState(1).permutation(n).astype(np.int32) if leaf_order else None
    for it in range(5):                                            # late iterations: rows with 1-3 records (short sorts)
        newc, oldc = orc.new_build_candidates(heap, k, rng_state, 8)
        e_idx, e_dist, e_flag = (a.copy() for a in heap)
        c_emu = emu_lib.emu_process_candidates(np.ascontiguousarray(X), n, d, 0, None, e_idx, e_dist, e_flag, k, newc, oldc, k,
                                               None if order is None else order.ctypes.data_as(C.c_void_p), kernel, 32, 3, 4, rank_scatter, short_sort)
        c_orc, _ = orc.process_candidates(X, "euclidean", heap, newc, oldc, 8)
        assert c_emu == c_orc and (c_orc > 0 or it >= 3)
        assert np.array_equal(e_idx, heap[0]) and np.array_equal(e_dist, heap[1]) and np.array_equal(e_flag, heap[2])


@pytest.mark.parametrize("split", [0, 1])
def test_leaf_init_fused_and_split(emu_lib, split):
    """init_rp_tree (pynndescent_.py:116-185): the fused leaf kernel (k_join mode 2 -- the hardware-verified path, so this
    validates the emulator) and the NND_LEAF_SPLIT pair (k_join mode 3 writes the distance matrices, k_leaf_push replays the
    pushes) must both leave exactly the oracle's heaps: indices, distance bits, flags, layout -- over a 4-tree forest, i.e.
    on heaps that are empty for the first tree and partly filled afterwards."""
    from sklearn.utils import check_random_state
    from oracle import nnd_oracle as orc
    n, d, k = 800, 20, 10
    X = latent(n, d, 31)
    rs = check_random_state(9)
    la, _ = orc.make_forest(X, k, 4, 24, rs, False, 200, "level")        # leaf_size 24: ~100 leaves per tree, sizes 2..24
    heap = orc.make_heap(n, k)
    orc.init_rp_tree(X, "euclidean", heap, la)
    sizes = (la >= 0).sum(1)
    assert la.shape[0] > 100 and sizes.min() >= 1 and sizes.max() <= 24
    pts = la[la >= 0].astype(np.int32)
    offs = np.zeros(la.shape[0] + 1, np.int64); np.cumsum(sizes, out=offs[1:])
    assert offs[-1] == 4 * n
    e = orc.make_heap(n, k)
    rc = emu_lib.emu_leaf_init(np.ascontiguousarray(X), n, d, d, pts, offs, la.shape[0], int(sizes.max()), e[0], e[1], e[2], k, split)
    assert rc == 0
    assert np.array_equal(e[0], heap[0]) and np.array_equal(e[1], heap[1]) and np.array_equal(e[2], heap[2])


@pytest.mark.parametrize("metric,mid,param", [("inner_product", 17, None), ("minkowski", 18, 3.0), ("minkowski", 18, 1.5)])
def test_generic_join_and_pair_distances_for_the_metrics_added_last(emu_lib, metric, mid, param):
    """inner_product (the dot class's second tail, picked by the RUNTIME metric id) and minkowski (p and 1/p from the first 16
    bytes of aux; x ** p by the deterministic exp2(p log2 x) restated in the oracle) through the source of the generic k_join
    (fused leaf init: distance phase, filter, per-member pushes) and of pair_distance_seq: heaps and distances equal the oracle's
    bit for bit."""
    from sklearn.utils import check_random_state
    from oracle import nnd_oracle as orc
    n, d, k = 500, 20, 8
    X = np.ascontiguousarray(latent(n, d, 37))
    X[7] = X[8]; X[9] = 0.0
    orc.set_metric_param(metric, None if param is None else {"p": param})
    aux = np.zeros(2, np.float64)
    if param is not None:
        aux[:] = (param, 1.0 / param)
    auxp = aux.ctypes.data_as(C.c_void_p)
    rs = np.random.RandomState(1)
    a = rs.randint(0, n, 3000).astype(np.int32); b = rs.randint(0, n, 3000).astype(np.int32)
    got = np.zeros(3000, np.float32)
    emu_lib.emu_pair_distances(X, d, d, mid, auxp, a, b, 3000, got)
    want = np.array([orc.dist(metric, X[i], X[j]) for i, j in zip(a, b)], np.float32)
    assert np.array_equal(got, want)
    la, _ = orc.make_forest(X, k, 3, 20, check_random_state(5), False, 200, "level")
    heap = orc.make_heap(n, k)
    orc.init_rp_tree(X, metric, heap, la)
    sizes = (la >= 0).sum(1)
    pts = la[la >= 0].astype(np.int32)
    offs = np.zeros(la.shape[0] + 1, np.int64); np.cumsum(sizes, out=offs[1:])
    e = orc.make_heap(n, k)
    rc = emu_lib.emu_leaf_init_metric(X, n, d, d, mid, auxp, pts, offs, la.shape[0], int(sizes.max()), e[0], e[1], e[2], k)
    assert rc == 0
    assert np.array_equal(e[0], heap[0]) and np.array_equal(e[1], heap[1]) and np.array_equal(e[2], heap[2])
    if metric == "inner_product":
        assert ((heap[1] > 1e30) & np.isfinite(heap[1])).any()                      # FLOAT32_MAX sentinels took part in the pushes


# ---- RP forest (k_forest.cu): every kernel of the level-synchronous forest under emulation against the oracle's level scheme
# (the GPU suite's test_forest_bit_exact_vs_oracle_level_scheme, on the CPU), and the point-major hyperplane pass NND_FOREST_PM
@pytest.mark.parametrize("n,d,angular,trees,leaf,pm,pm_limit", [
    (200, 20, 0, 3, 16, 0, 0),                 # the hardware-verified position-major pass (k_forest_side128) validates the harness
    (200, 20, 0, 3, 16, 1, 500),               # point-major while the level's table is <= 500 B (6 nodes), position-major below
    (200, 128, 1, 2, 12, 1, 1 << 30),          # point-major on every level; angular trees, full 128-float rows (every lane holds data)
    (160, 7, 1, 2, 10, 1, 1 << 30),            # padded rows (ld = 8: lanes 2..31 hold nothing)
    (90, 16, 0, 12, 10, 1, 1 << 30),           # the default forest size: 12 trees = 12 lanes look positions up
    (120, 130, 0, 2, 10, 1, 1 << 30),          # ld = 132: two-chunk rows, only lane 0 holds data of the second chunk
    (100, 300, 1, 2, 10, 1, 1 << 30),          # ld = 300: four-chunk instantiation, angular
    (90, 784, 0, 2, 10, 1, 10000),             # ld = 784 (cfg5 rows): eight-chunk instantiation on the shallow levels, k_forest_side below
    (100, 1030, 0, 2, 10, 1, 1 << 30),         # ld > 1024: the switch must leave the generic k_forest_side in charge
])  # (every emulated CUDA thread is an OS thread: sizes are kept small; 700 x 20 x 5 trees and 500 x 128 angular passed as well)
def test_forest_kernels_and_point_major_pass(emu_lib, n, d, angular, trees, leaf, pm, pm_limit):
    from oracle import nnd_oracle as orc
    X = latent(n, d, 17 + d)
    if angular:
        X[3] = 0.0                                                            # a zero row: norm guard of the angular hyperplane
    X[5] = X[6]                                                               # duplicates: margin ties -> hash coin
    ld = (d + 3) // 4 * 4
    Xp = np.zeros((n, ld), np.float32); Xp[:, :d] = X
    states = np.random.RandomState(3).randint(-2 ** 31 + 1, 2 ** 31 - 1, size=(trees, 3)).astype(np.int64)
    lp = np.full(trees * n, -7, np.int32); lo = np.zeros(trees * n + 1, np.int64); info = np.zeros(8, np.int64)
    rc = emu_lib.emu_forest(Xp, n, ld, angular, states, trees, leaf, 200, pm, pm_limit, lp, lo, info)
    assert rc == 0
    nl, ml, levels, levels_pm = (int(v) for v in info[:4])
    want = orc.forest_leaf_array(X, states, leaf, bool(angular), 200, "level")
    assert nl == want.shape[0] and ml == want.shape[1]
    got = np.full_like(want, -1)
    for i in range(nl):
        got[i, : lo[i + 1] - lo[i]] = lp[lo[i]: lo[i + 1]]
    assert np.array_equal(got, want)
    if not pm or ld > 1024:
        assert levels_pm == 0
    elif pm_limit >= 1 << 30:
        assert levels_pm == levels > 3
    else:
        assert 0 < levels_pm < levels                                         # both kernels took part in this forest


# ---- query graph walk (k_query.cu) under emulation: the reference's own query() answers on its prepared structures, and the
# oracle's search bit for bit -- also for the metrics added after the last hardware run
def _pad(a, ld):
    out = np.zeros((a.shape[0], ld), np.float32); out[:, : a.shape[1]] = a
    return out


def _query_aux(emu_lib, metric, data_p, d):
    from oracle import nnd_oracle as orc
    n = data_p.shape[0]
    if metric == "cosine":
        aux = np.zeros(n, np.float32); emu_lib.emu_norms(data_p, n, data_p.shape[1], aux)
        return aux
    if metric == "jaccard":
        return (data_p != 0).sum(1).astype(np.float32)
    if metric == "minkowski":
        return None
    return np.zeros(4, np.float32)


def _emu_query(emu_lib, G, metric, mid, Q, k, eps, aux_override=None, n_neighbors=None):
    d = G["data"].shape[1]; ld = (d + 3) // 4 * 4
    data_p = _pad(np.asarray(G["data"], np.float32), ld)
    aux = aux_override if aux_override is not None else _query_aux(emu_lib, metric, data_p, d)
    hi = np.full((Q.shape[0], k), -7, np.int32); hd = np.zeros((Q.shape[0], k), np.float32); st = np.zeros(4, np.int64)
    rc = emu_lib.emu_query(data_p, data_p.shape[0], d, ld, mid, aux.ctypes.data_as(C.c_void_p), np.ascontiguousarray(G["indptr"], np.int32),
                           np.ascontiguousarray(G["indices"], np.int32), _pad(np.asarray(G["hyperplanes"], np.float32)[:, :d], ld),
                           np.ascontiguousarray(G["offsets"], np.float32), np.ascontiguousarray(G["children"], np.int32),
                           np.ascontiguousarray(G["tree_indices"], np.int32), _pad(np.asarray(Q, np.float32), ld), Q.shape[0], k, eps,
                           float(G["min_distance"]), int(G["n_neighbors"]) if n_neighbors is None else n_neighbors, int(metric in ("cosine", "dot")),
                           np.ascontiguousarray(G["search_rng_state"], np.int64), hi, hd, st)
    assert rc == 0
    return hi, hd, st


@pytest.mark.parametrize("name,mid", [("query_euclid_3k32", 0), ("query_cosine_2k24", 1), ("query_manhattan_2k20", 4), ("query_jaccard_2k40", 8)])
def test_query_walk_matches_reference_answers(emu_lib, name, mid):
    from conftest import load_golden
    from oracle import nnd_oracle as orc
    G = load_golden(name); metric = str(G["metric"])
    Q = G["Q"][:48]
    k, eps = int(G["k_query"]), float(G["epsilon"])
    hi, hd, st = _emu_query(emu_lib, G, metric, mid, Q, k, eps)
    oi, od = orc.search(G["data"], metric, G["indptr"], G["indices"], G["hyperplanes"], G["offsets"], G["children"], G["tree_indices"], Q, k,
                        eps, float(G["min_distance"]), int(G["n_neighbors"]), G["search_rng_state"])
    assert np.array_equal(hi, oi) and np.array_equal(hd, od)                    # raw result heaps: the oracle's, bit for bit
    si, sd = orc.deheap_sort(hi.copy(), hd.copy())
    assert np.array_equal(G["vertex_order"][si], G["out_idx"][:48])             # ... and the reference's query() answers
    np.testing.assert_allclose(orc.correction(metric, sd), G["out_dist"][:48], rtol=1e-5, atol=1e-6)
    assert st[0] > 0


@pytest.mark.parametrize("metric,mid,param", [("inner_product", 17, None), ("minkowski", 18, 3.0)])
def test_query_walk_for_the_metrics_added_last(emu_lib, metric, mid, param):
    """The euclidean golden's search structures walked under inner_product / minkowski: whatever the graph was built for, kernel and
    oracle must agree on every visited vertex and distance (the structures are inputs; the metric only enters query_distance)."""
    from conftest import load_golden
    from oracle import nnd_oracle as orc
    G = load_golden("query_euclid_3k32")
    orc.set_metric_param(metric, None if param is None else {"p": param})
    aux = np.zeros(2, np.float64)
    if param is not None:
        aux[:] = (param, 1.0 / param)
    Q = G["Q"][:32]
    hi, hd, st = _emu_query(emu_lib, G, metric, mid, Q, 10, 0.15, aux_override=aux)
    oi, od = orc.search(G["data"], metric, G["indptr"], G["indices"], G["hyperplanes"], G["offsets"], G["children"], G["tree_indices"], Q, 10,
                        0.15, float(G["min_distance"]), int(G["n_neighbors"]), G["search_rng_state"])
    assert np.array_equal(hi, oi) and np.array_equal(hd, od)


def test_query_walk_rng_replay_pass(emu_lib):
    """min(k, n_neighbors) = 40 is larger than any search-tree leaf (<= 30 points): every query needs random seeds (pynndescent_.py
    :1835-1847), i.e. the shared rng stream -- pass 1 flags them all, pass 2
    (k_query_seq: one warp, query order, global seed heap and visited bitmap) replays them with the real stream; a zero query under
    cosine stays unfilled.  Raw heaps equal the oracle's."""
    from conftest import load_golden
    from oracle import nnd_oracle as orc
    G = load_golden("query_cosine_2k24")
    Q = G["Q"][:12].copy()
    Q[3] = 0.0
    hi, hd, st = _emu_query(emu_lib, G, "cosine", 1, Q, 40, 0.2, n_neighbors=40)
    oi, od = orc.search(G["data"], "cosine", G["indptr"], G["indices"], G["hyperplanes"], G["offsets"], G["children"], G["tree_indices"], Q, 40,
                        0.2, float(G["min_distance"]), 40, G["search_rng_state"])
    assert st[2] >= 11                                                          # (almost) every query went through the replay pass
    keep = np.arange(12) != 3
    assert np.array_equal(hi[keep], oi[keep]) and np.array_equal(hd[keep], od[keep])
    assert np.all(oi[3] == -1)


# ---- init_random, deheap_sort, diversify (the kernels around the descent loop that evaluate pair distances one by one)
@pytest.mark.parametrize("metric,mid,param", [("euclidean", 0, None), ("inner_product", 17, None), ("minkowski", 18, 3.0)])
def test_init_random_deheap_sort_and_diversify(emu_lib, metric, mid, param):
    """init_random (pynndescent_.py:188-203: rows the leaves left short draw from the ONE sequential rng stream), deheap_sort
    (utils.py:189-218) and the forward diversification (pynndescent_.py:369-403) under emulation: heaps, rng state, sorted graph and
    pruned rows equal the oracle's bit for bit."""
    from sklearn.utils import check_random_state
    from oracle import nnd_oracle as orc
    n, d, k = 400, 20, 12
    X = np.ascontiguousarray(latent(n, d, 41))
    orc.set_metric_param(metric, None if param is None else {"p": param})
    aux = np.zeros(2, np.float64)
    if param is not None:
        aux[:] = (param, 1.0 / param)
    auxp = aux.ctypes.data_as(C.c_void_p)
    rs = check_random_state(11)
    rng0 = rs.randint(orc.INT32_MIN, orc.INT32_MAX, 3).astype(np.int64)
    la, _ = orc.make_forest(X, k, 2, 8, rs, False, 200, "level")          # tiny leaves: most rows stay short after the leaf init
    heap = orc.make_heap(n, k)
    orc.init_rp_tree(X, metric, heap, la)
    e = tuple(a.copy() for a in heap)
    rng_o, rng_e = rng0.copy(), rng0.copy()
    orc.init_random(X, metric, heap, rng_o)
    total = emu_lib.emu_init_random(X, n, d, d, mid, auxp, e[0], e[1], e[2], k, rng_e)
    assert total > n                                                          # more than one draw per row on average
    assert np.array_equal(rng_e, rng_o)
    assert np.array_equal(e[0], heap[0]) and np.array_equal(e[1], heap[1]) and np.array_equal(e[2], heap[2])
    si, sd = orc.deheap_sort(heap[0].copy(), heap[1].copy())
    out = np.zeros_like(e[1])
    emu_lib.emu_deheap_sort(e[0], e[1], n, k, 0, out)
    assert np.array_equal(e[0], si) and np.array_equal(out, sd) and np.array_equal(e[1], sd)
    oi, od = si.copy(), sd.copy()
    orc.lib().orc_diversify(oi, od, n, k, X, d, orc.METRIC_IDS[metric])
    gi, gd = si.copy(), sd.copy()
    emu_lib.emu_diversify(X, n, d, d, mid, auxp, gi, gd, k)
    assert (oi < 0).sum() > (si < 0).sum()                                   # the pass pruned something
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)


# ---- hub search tree (k_prepare.cu k_hub_*) + the host flattening: the REFERENCE's flat search tree from its own kNN graph
@pytest.mark.skipif(os.environ.get("NND_SLOW_EMU") != "1", reason="2.5-4 minutes per golden (k_hub_side visits all N points on every one "
                    "of ~40 levels, one OS thread per emulated CUDA thread); run with NND_SLOW_EMU=1 -- all three passed when written")
@pytest.mark.parametrize("name", ["query_euclid_3k32", "query_cosine_2k24", "query_jaccard_2k40"])
def test_hub_tree_kernels_reproduce_the_reference_search_tree(emu_lib, name):
    """make_hub_tree (rp_trees.py:714-1309) level by level under emulation, then prepare.flatten_hub_tree: hyperplanes, offsets,
    children and leaf order are the reference's own arrays (goldens of make_golden_query.py) -- euclidean splits, angular splits,
    and binary data under angular splits, where most margins are ties."""
    from conftest import load_golden
    from oracle import nnd_oracle as orc
    G = load_golden(name); metric = str(G["metric"])
    X = np.asarray(G["X"], np.float32)
    n, d = X.shape; ld = (d + 3) // 4 * 4
    Xp = _pad(X, ld)
    ng = np.ascontiguousarray(G["ng_idx"], np.int32)
    ncap = 2 * n + 16; hyper_cap = n // 8 + 64
    indices = np.zeros(n, np.int32)
    ns, nz, nl, nr, nh = (np.zeros(ncap, np.int32) for _ in range(5))
    hyp = np.zeros((hyper_cap, ld), np.float32); off = np.zeros(hyper_cap, np.float32); counts = np.zeros(4, np.int32)
    rc = emu_lib.emu_hub_tree(Xp, n, ld, int(metric in orc.ANGULAR_METRICS), ng, ng.shape[1], 30, 200, indices, ns, nz, nl, nr, nh, hyp, off,
                              hyper_cap, counts)
    assert rc == 0 and counts[2] == 0
    n_nodes, n_hyper = int(counts[0]), int(counts[1])
    assert np.array_equal(indices, G["vertex_order"])                            # leaf order = the reference's tree.indices
    from pynndescent_b200.prepare import flatten_hub_tree                        # the product's host-side flattening
    hyperplanes, offsets, children = flatten_hub_tree(ns[:n_nodes], nz[:n_nodes], nl[:n_nodes], nr[:n_nodes], nh[:n_nodes], hyp[:n_hyper, :d],
                                                      off[:n_hyper], d)
    assert np.array_equal(children, G["children"])
    np.testing.assert_allclose(hyperplanes, G["hyperplanes"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(offsets, G["offsets"], rtol=1e-5, atol=1e-6)
