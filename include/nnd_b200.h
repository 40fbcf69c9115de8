/*
 * nnd_b200.h -- C-ABI of libnnd_b200.so: the B200 (sm_100a) NN-descent kNN-graph hot path.
 *
 * This is the drop-in boundary for lmcinnes/pynndescent's numba seams:
 *   seam 1  make_forest + rptree_leaf_array      pynndescent/pynndescent_.py:1118-1130
 *   seam 2  nn_descent(...) -> (indices, dists)   pynndescent/pynndescent_.py:1247-1260 (def :324-366)
 *   seam 3  _search_function / _deheap_function    pynndescent/pynndescent_.py:2336-2361   (next round)
 *
 * Conventions: extern "C"; every call returns 0 on success or a negative NND_ERR_* code and
 * the text is available from nnd_last_error(); no exceptions cross the boundary.  All array
 * arguments are DEVICE pointers unless the name ends in _host.  The caller owns every buffer
 * (the Python host takes them from torch.Tensor.data_ptr()); scratch comes from an explicit
 * workspace whose size is queried first -- the library never allocates device memory behind
 * the caller's back except the small per-context constant tables made in nnd_create().
 * `stream` is a cudaStream_t passed as void*.  A context is bound to one device and is not
 * thread-safe.
 *
 * Parity contract: given the same inputs, every stage reproduces the
 * reference's result bit-for-bit (indices, flags, heap layout, accepted-push count `c`) for a
 * declared emulated numba thread count T; distances follow one defined summation order and
 * agree with the reference's fastmath numbers to <= 1e-5 relative.
 */
#ifndef NND_B200_H
#define NND_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define NND_API __attribute__((visibility("default")))
#else
#define NND_API
#endif

#define NND_OK 0
#define NND_ERR_INVALID (-1)     /* bad argument                                     */
#define NND_ERR_CUDA (-2)        /* a CUDA runtime call failed                        */
#define NND_ERR_WORKSPACE (-3)   /* workspace too small                               */
#define NND_ERR_UNSUPPORTED (-4) /* metric / shape not implemented on the GPU path    */
#define NND_ERR_OVERFLOW (-5)    /* an internal list would overflow (never silent)    */

/* distances.py:2103-2188 names -> ids ("fast alternatives" are what the path computes in;
 * the correction back to the true distance happens in nnd_deheap_sort).                 */
#define NND_METRIC_SQEUCLIDEAN 0 /* distances.py:63-91   (euclidean / l2 / sqeuclidean) */
#define NND_METRIC_ALT_COSINE 1  /* distances.py:583-630 (cosine)                       */
#define NND_METRIC_ALT_DOT 2     /* distances.py:665-701 (dot; data pre-normalised)     */
#define NND_METRIC_CORRELATION 3 /* distances.py:1283-1320                              */
/* SURVEY 8f-2 "more named metrics": computed by the generic (unbatched) join, the query walk and prepare() */
#define NND_METRIC_MANHATTAN 4   /* distances.py manhattan / taxicab / l1   (f64 accumulator like the reference) */
#define NND_METRIC_CHEBYSHEV 5   /* distances.py chebyshev / linfinity / linfty / linf                           */
#define NND_METRIC_CANBERRA 6    /* distances.py canberra                   (f64 accumulator)                     */
#define NND_METRIC_HAMMING 7     /* distances.py hamming                                                         */
/* binary family: one accumulator tt = #{x_i != 0 and y_i != 0} + the per-row non-zero counts (metric aux) */
#define NND_METRIC_ALT_JACCARD 8 /* distances.py:303-326 alternative_jaccard (jaccard; correction 1 - 2^-d)       */
#define NND_METRIC_DICE 9
#define NND_METRIC_MATCHING 10
#define NND_METRIC_KULSINSKI 11
#define NND_METRIC_ROGERSTANIMOTO 12 /* also sokalmichener (same formula in the reference)                       */
#define NND_METRIC_RUSSELLRAO 13
#define NND_METRIC_SOKALSNEATH 14
#define NND_METRIC_YULE 15
#define NND_METRIC_ALT_HELLINGER 16 /* distances.py:1387-1416 alternative_hellinger (correction sqrt(1 - 2^-d)) */
/* distances.py:759-790 alternative_inner_product (inner_product): 1 / <x, y>, FLOAT32_MAX when <x, y> <= 0; the same f32 dot
 * accumulation as NND_METRIC_ALT_DOT (one kernel class), no normalisation, correction -1 / d (distances.py:842-853) */
#define NND_METRIC_ALT_INNER 17
/* distances.py:138-153 minkowski(x, y, p): float64 result += (f32 |x_i - y_i|) ** p, then result ** (1 / p).  p reaches the
 * reference's function positionally from metric_kwds.values() (pynndescent_.py:1064-1065, 1294-1302); here the caller stores it
 * in the first 16 bytes of the metric aux buffer with nnd_metric_set_param (two doubles: p, 1 / p).  x ** p is evaluated by a
 * deterministic exp2(p * log2 x) (fma chains only), restated in the oracle: f32 results equal the reference's on its KATs */
#define NND_METRIC_MINKOWSKI 18
#define NND_METRIC_MAX 18

#define NND_CORRECT_NONE 0
#define NND_CORRECT_SQRT 1   /* distances.py:2171       np.sqrt                         */
#define NND_CORRECT_ALTCOS 2 /* distances.py:704-711    1 - 2^-d                        */

typedef struct nnd_ctx nnd_ctx;

NND_API const char *nnd_last_error(void);
NND_API int nnd_version(void);

/* context: device-resident jump-ahead tables for the Tausworthe rng (utils.py:17-57) */
NND_API int nnd_create(nnd_ctx **out, int device);
NND_API int nnd_destroy(nnd_ctx *ctx);

/* ---- problem description shared by all stage calls ---------------------------------- */
typedef struct {
    const float *X; /* [N, ld] row-major, ld % 4 == 0, columns d..ld-1 zero, 16 B aligned */
    int64_t N;
    int32_t d, ld;
    int32_t metric;  /* NND_METRIC_*                                                      */
    const void *aux; /* from nnd_metric_prepare: cosine f32 norm2[N]; correlation f64 (mu,nrm2)[N];
                        binary family f32 nnz[N]; hellinger f32 l1[N]; nothing for the other metrics   */
    int32_t k;       /* n_neighbors                                                        */
    int32_t *heap_idx;  /* [N,k]  utils.py:130-158 make_heap layout                        */
    float *heap_dist;   /* [N,k]                                                           */
    uint8_t *heap_flag; /* [N,k]                                                           */
} nnd_problem;

/* bytes of `aux` for (metric, N) */
NND_API int nnd_metric_aux_bytes(int32_t metric, int64_t N, size_t *bytes);
/* fills aux (per-point squared norms / means) in the canonical summation order */
NND_API int nnd_metric_prepare(nnd_ctx *ctx, const float *X, int64_t N, int32_t d, int32_t ld, int32_t metric, void *aux,
                       void *stream);
/* parameter of a metric that takes one -- the reference passes metric_kwds.values() positionally to the distance function
 * (pynndescent_.py:1064-1065, 1294-1302); NND_METRIC_MINKOWSKI: p > 0 (distances.py:138).  Stored in the first 16 bytes of aux
 * (every aux buffer has them, nnd_metric_aux_bytes); call after nnd_metric_prepare, before any stage.                        */
NND_API int nnd_metric_set_param(nnd_ctx *ctx, int32_t metric, void *aux, double value, void *stream);
/* utils.py:130-158 make_heap: idx=-1, dist=+inf, flag=0 */
NND_API int nnd_make_heap(nnd_ctx *ctx, int32_t *heap_idx, float *heap_dist, uint8_t *heap_flag, int64_t N, int32_t k,
                  void *stream);

/* ---- seam 1: RP forest (rp_trees.py:2815-2922) --------------------------------------- */
/* Level-synchronous forest with the reference's split rule and stable partition; node
 * randomness = hash(tree seed, level, node ordinal).  Output is the
 * leaf set in CSR form: leaf_points[n_trees*N] (tree-major, leaves contiguous, DFS left-to-
 * right order, ascending point id inside a leaf) and leaf_offsets[n_leaves+1].            */
NND_API int nnd_forest_workspace_bytes(int64_t N, int32_t d, int32_t n_trees, size_t *bytes);
NND_API int nnd_forest_build(nnd_ctx *ctx, const float *X, int64_t N, int32_t d, int32_t ld, int32_t angular,
                     const int64_t *tree_rng_states_host /* [n_trees*3] rp_trees.py:2850-2852 */, int32_t n_trees,
                     int32_t leaf_size, int32_t max_depth, int32_t *leaf_points /* [n_trees*N] */,
                     int64_t *leaf_offsets /* [n_trees*N+1] capacity */, int64_t *n_leaves_host,
                     int32_t *max_leaf_host, void *ws, size_t ws_bytes, void *stream);

/* ---- seam 2: NN-descent stages (pynndescent_.py:324-366) ------------------------------ */
/* workspace for leaf init / descent iterations.  push_capacity = max half-updates one
 * routing round may hold (rows are processed in blocks sized so this is never exceeded).  */
NND_API int nnd_descent_workspace_bytes(int64_t N, int32_t k, int32_t C, int32_t T, int64_t push_capacity, size_t *bytes);

/* pynndescent_.py:116-185 init_rp_tree (+ :73-113 generate_leaf_updates) from CSR leaves.
 * n_updates_host (optional) receives the number of half-updates routed.                   */
NND_API int nnd_leaf_init(nnd_ctx *ctx, const nnd_problem *P, const int32_t *leaf_points, const int64_t *leaf_offsets,
                  const int64_t *leaf_offsets_host, int64_t n_leaves, int32_t max_leaf, int64_t push_capacity,
                  void *ws, size_t ws_bytes, int64_t *n_pushes_host, void *stream);

/* pynndescent_.py:188-203 init_random; rng_state_host[3] is advanced exactly as the reference
 * advances it.                                                                             */
NND_API int nnd_init_random(nnd_ctx *ctx, const nnd_problem *P, int64_t *rng_state_host, void *ws, size_t ws_bytes,
                    void *stream);

/* utils.py:221-320 new_build_candidates, emulating T numba threads (stream n = rng_state+n,
 * vertex ranges of N//T+1).  Outputs new_cand/old_cand int32[N,C] with the reference's heap
 * layout; clears heap_flag of sampled new entries.                                          */
NND_API int nnd_cand_workspace_bytes(int64_t N, int32_t k, int32_t C, int32_t T, size_t *bytes);
/* tables_valid != 0: the T priority tables at the head of `ws` were filled by an earlier call with
 * the same (rng_state, T, N, k) and `ws` was not touched since (the streams never advance).     */
NND_API int nnd_build_candidates(nnd_ctx *ctx, const nnd_problem *P, int32_t C, const int64_t *rng_state_host, int32_t T,
                         int32_t tables_valid, int32_t *new_cand, int32_t *old_cand, void *ws, size_t ws_bytes,
                         void *stream);

/* pynndescent_.py:217-263 process_candidates == utils.py:536-658 generate_graph_update_array
 * + utils.py:661-733 apply_graph_update_array over all rows.  c_host = accepted pushes.
 * stats_host (optional, int64[4]): gathered rows, evaluated pairs, routed half-updates, blocks */
NND_API int nnd_process_candidates(nnd_ctx *ctx, const nnd_problem *P, int32_t C, const int32_t *new_cand,
                           const int32_t *old_cand, int64_t push_capacity, void *ws, size_t ws_bytes,
                           int64_t *c_host, int64_t *stats_host, void *stream);

/* utils.py:189-218 deheap_sort (in place, exact tie order) fused with the distance
 * correction of neighbor_graph (pynndescent_.py:2145-2158): out_dist may alias heap_dist.  */
NND_API int nnd_deheap_sort(nnd_ctx *ctx, int32_t *idx, float *dist, int64_t N, int32_t k, int32_t correction,
                    float *out_dist, void *stream);

/* the whole of nn_descent_internal (pynndescent_.py:266-320): iterate build_candidates +
 * process_candidates until c <= delta*k*N or n_iters.  c_hist_host[n_iters] optional.      */
NND_API int nnd_descent(nnd_ctx *ctx, const nnd_problem *P, int32_t C, const int64_t *rng_state_host, int32_t T,
                int32_t n_iters, double delta, int64_t push_capacity, void *ws, size_t ws_bytes,
                int32_t *iters_run_host, int64_t *c_hist_host, int64_t *stats_host, void *stream);

/* Schedule switches read from the environment by nnd_create (none changes a result; csrc/nnd_common.cuh, DESIGN.md section 10a;
 * defaults = the winners of the round-2 hardware A/B, profiles/ab_r02_1m_switches.json): NND_RANK_SCATTER, NND_LEAF_SPLIT,
 * NND_SMEM_CARVEOUT, NND_FOREST_PM (+ NND_FOREST_PM_MB), NND_CAND_SKIP, NND_APPLY_SHORT_SORT, NND_CAND_PARTITION.           */

/* ---- sharded build (one process per GPU) ----------------------------------------------
 * Every process holds X and the whole heap; it samples candidates and joins only rows [lo,hi), exchanges the
 * half-updates with the owners of their target rows and applies what it receives.                            */
NND_API int nnd_set_row_range(nnd_ctx *ctx, int64_t lo, int64_t hi);
/* the local join filters with thr[v] = heap_dist[v, 0] of ALL vertices (utils.py:633, pynndescent_.py:246); a process that owns
 * only a row range of the heap supplies the snapshot itself (N floats in device memory, all-gathered from the owners; NULL =
 * snapshot the local heap as on one GPU).  Stale / larger thresholds only admit a superset of pushes (utils.py:484 re-checks).  */
NND_API int nnd_set_thresholds(nnd_ctx *ctx, const float *thr);
/* The local join and the replay drop half-updates (t <- o, d) whose o already sits in heap row t, and all but the first of
 * several with the same (t, o): they can never be accepted -- PROVIDED every distance stored in the heap is the metric's own
 * value for its pair (then an evicted o was the row's maximum with exactly this d, so `d >= root` rejects it, utils.py:484).
 * A warm start with caller-supplied init_dist (pynndescent_.py:1225-1242 pushes them verbatim) breaks that premise:
 * trusted = 0 switches both filters off for this context (default 1).                                                  */
NND_API int nnd_set_heap_trusted(nnd_ctx *ctx, int32_t trusted);
/* new_build_candidates (utils.py:221-320) for the own rows only, in three steps around two collectives of the caller:
 *  1. nnd_cand_local_counts: out_host[5T + 1 + R] = for every stream n this process's partial prefixes {prR_n(s_n), prV(s_n),
 *     prV(e_n), prR_n(e_n), total R_n}, then the total of valid edges, then the number of reverse-edge records per owner process;
 *  2. (caller: all-gather of the tables) nnd_cand_export: glob_host[6T] = per stream {R_s, V_s, V_e, R_e summed over all
 *     processes, sum of the lower ranks' R_n totals, sum of the lower ranks' valid totals}; forward priorities stay local, one
 *     12-byte record {target, scan key, priority} per reverse edge is appended to records_out, grouped by the target's owner
 *     process in rank order (step 1's per-owner counts are the segment sizes);
 *  3. (caller: all-to-all-v of the buckets) nnd_cand_import_select: received records -> candidate lists of the own rows, sampled
 *     flags cleared.  Identical lists to nnd_build_candidates for any number of processes.                               */
NND_API int nnd_cand_local_counts(nnd_ctx *ctx, const nnd_problem *P, int32_t C, int32_t T, int32_t R, int64_t owner_block, void *ws,
                          size_t ws_bytes, int64_t *out_host, void *stream);
NND_API int nnd_cand_export(nnd_ctx *ctx, const nnd_problem *P, int32_t C, const int64_t *rng_state_host, int32_t T, int32_t tables_valid,
                    int32_t R, int64_t owner_block, const int64_t *glob_host, void *records_out, void *ws, size_t ws_bytes,
                    void *stream);
NND_API int nnd_cand_import_select(nnd_ctx *ctx, const nnd_problem *P, int32_t C, int32_t T, const void *records, int64_t n,
                           int32_t *new_cand, int32_t *old_cand, void *ws, size_t ws_bytes, void *stream);
/* prepass of one process_candidates call restricted to the row range; n_rounds_host = routing rounds needed */
NND_API int nnd_join_prepare(nnd_ctx *ctx, const nnd_problem *P, int32_t C, const int32_t *new_cand, const int32_t *old_cand,
                             int64_t push_capacity, void *ws, size_t ws_bytes, int32_t *n_rounds_host, void *stream);
/* run the join of one round; the half-updates stay in the workspace (rounds beyond the plan emit nothing).
 * Rounds are planned optimistically: if this one does not fit push_capacity it is split in two (the plan grows
 * by one round), *n_pushed_host = -1 and the caller emits the same round index again.  NND_ERR_OVERFLOW only
 * when a single unit does not fit.                                                                             */
NND_API int nnd_join_emit_round(nnd_ctx *ctx, const nnd_problem *P, int32_t C, int64_t push_capacity, void *ws, size_t ws_bytes,
                                int32_t round, int64_t *n_pushed_host, void *stream);
/* group them by owner process (owner = target / owner_block) into 24-byte records {tgt i32, oth i32, d f32, pad, seq u64} */
NND_API int nnd_push_export(nnd_ctx *ctx, const nnd_problem *P, int32_t C, int64_t push_capacity, void *ws, size_t ws_bytes,
                            int64_t n_pushed, int64_t owner_block, int32_t R, void *out, int64_t *counts_host, void *stream);
/* apply received records (any order) to the heap rows of this process's row range (nnd_set_row_range; all rows by
 * default): counting sort by target, then checked_flagged_heap_push (utils.py:471-533) replayed per row in ascending
 * `seq`; c_host = accepted pushes.  A record outside the row range is an error.  Also the warm start of
 * pynndescent_.py:1225-1242 / utils.py:836-860 (records = the init graph's entries, seq = column position).          */
NND_API int nnd_apply_records(nnd_ctx *ctx, const nnd_problem *P, int32_t C, const void *records, int64_t n,
                              int64_t push_capacity, void *ws, size_t ws_bytes, int64_t *c_host, void *stream);
/* fused leaf join + apply on leaves [l0,l1) of one tree (disjoint point sets) */
NND_API int nnd_leaf_fused_range(nnd_ctx *ctx, const nnd_problem *P, const int32_t *leaf_points, const int64_t *leaf_offsets,
                                 int64_t l0, int64_t l1, int32_t max_leaf, void *stream);
/* the same, through the split leaf init (distance-matrix kernel + thread-per-member push kernel; NND_LEAF_SPLIT) when that is
 * enabled and its scratch fits the routing workspace `ws`, else through the fused kernel.  pos0 / npos = position range of the
 * leaves [l0,l1) in leaf_points (leaf_offsets[l0], leaf_offsets[l1] - leaf_offsets[l0]; the caller has them on the host).  */
NND_API int nnd_leaf_range(nnd_ctx *ctx, const nnd_problem *P, const int32_t *leaf_points, const int64_t *leaf_offsets, int64_t l0,
                           int64_t l1, int64_t pos0, int64_t npos, int32_t max_leaf, int64_t push_capacity, void *ws, size_t ws_bytes,
                           void *stream);
/* pack (unpack=0) / unpack (unpack=1) the heap rows listed in `rows` (NULL = rows 0..n-1) */
NND_API int nnd_heap_pack_rows(nnd_ctx *ctx, const nnd_problem *P, const int32_t *rows, int64_t n, int32_t *out_idx,
                               float *out_dist, uint8_t *out_flag, int32_t unpack, void *stream);

/* ---- seam 3: query graph walk (pynndescent_.py:1770-1883 search_closure) ------------------------------------
 * Prepared search structures in the reference's own format (pynndescent_.py:1664-1750): data re-ordered to
 * tree-leaf order (:1629-1641), CSR search graph (:1746-1747), flat search tree (rp_trees.py:2926-3049: leaves
 * encode -start/-end ranges of tree_indices in `children`).                                                    */
typedef struct {
    const float *data;   /* [N, ld] re-ordered, padded like nnd_problem.X           */
    int64_t N; int32_t d, ld; int32_t metric; const void *aux; /* aux of the re-ordered data (nnd_metric_prepare) */
    const int32_t *indptr, *indices;                      /* CSR search graph                                   */
    const float *hyperplanes;                             /* [n_nodes, ld]                                       */
    const float *offsets; const int32_t *children;        /* [n_nodes], [n_nodes, 2]                             */
    const int32_t *tree_indices;                          /* [N]                                                 */
    float min_distance; int32_t n_neighbors; int32_t normalize; /* :1539, :1748, :1762-1768 (cosine / dot)       */
} nnd_search_index;
NND_API int nnd_query_workspace_bytes(int64_t N, int64_t nq, int32_t k, size_t *bytes);
/* Q: device [nq, ld] zero-padded.  out_idx/out_dist: raw result heaps [nq,k] (apply nnd_deheap_sort next,
 * then vertex_order and the distance correction, :2361-2377).  rng_state_host = search_rng_state (:1796 copy).
 * stats_host[3] (optional): distance evaluations, seed pops, queries replayed sequentially for the rng.        */
NND_API int nnd_query(nnd_ctx *ctx, const nnd_search_index *S, const float *Q, int64_t nq, int32_t k, double epsilon,
                      int64_t *rng_state_host, int32_t *out_idx, float *out_dist, void *ws, size_t ws_bytes,
                      int64_t *stats_host, void *stream);

/* ---- search-graph preparation, GPU parts (pynndescent_.py:1333-1662) ------------------------------------------
 * hub tree (rp_trees.py:714-1309 make_hub_tree): node table of capacity 2N (node 0 = root; left/right = -1 for a
 * leaf whose points are indices[start:start+size]); hyperplane of an internal node = node_hyper[node_hyper_row[n]].
 * The host flattens this to the reference's pre-order arrays (rp_trees.py:2926-3049).                            */
NND_API int nnd_hub_tree_workspace_bytes(int64_t N, int32_t d, size_t *bytes);
NND_API int nnd_hub_tree_build(nnd_ctx *ctx, const float *X, int64_t N, int32_t d, int32_t ld, int32_t angular,
                               const int32_t *ng_idx, int32_t k, int32_t leaf_size, int32_t max_depth, int32_t *indices,
                               int32_t *node_start, int32_t *node_size, int32_t *node_left, int32_t *node_right,
                               int32_t *node_hyper_row, float *node_hyper, float *node_off, int64_t hyper_cap,
                               int32_t *n_nodes_host, int32_t *n_hyper_host, void *ws, size_t ws_bytes, void *stream);
/* pynndescent_.py:369-403 diversify (prune_probability 1) on a [N,k] neighbour table in ascending distance order
 * (in place, rows compacted, -1/+inf padded); applied to already-compacted rows it is diversify_csr :549-588.     */
NND_API int nnd_diversify(nnd_ctx *ctx, const nnd_problem *P, int32_t *ind, float *dist, void *stream);

/* The sparse-matrix part of _init_search_graph on the device (pynndescent_.py:1521-1651: G = max(G, G^T) of the twice
 * diversified graph, setdiag(0), degree_prune :728-760, != 0, rows and columns permuted to leaf order :1629-1634).
 * (di, dd) / (di2, dd2) = the rows after the first / second diversification pass (int32 / f32 [N,k], -1 holes anywhere;
 * each pass packs its own survivors, so the two pairs are NOT position-aligned; zero distances already FLOAT32_EPS, :1526),
 * vertex_order = the hub tree's leaf order.  Outputs the CSR of the search graph
 * (indptr int32[N+1], sorted indices; capacity 2*N*k entries is always enough), its size and min_distance (:1539).      */
NND_API int nnd_search_graph_workspace_bytes(int64_t N, int32_t k, size_t *bytes);
NND_API int nnd_search_graph_build(nnd_ctx *ctx, const int32_t *di, const float *dd, const int32_t *di2, const float *dd2, int64_t N, int32_t k,
                                   const int32_t *vertex_order, int32_t max_degree, int32_t *indptr_out, int32_t *indices_out,
                                   int64_t capacity, int64_t *nnz_host, float *min_distance_host, void *ws, size_t ws_bytes,
                                   void *stream);

/* ---- small exported primitives (unit tests / known-answer vectors) -------------------- */
/* utils.py:17-57: out[i] = i-th tau_rand_int (as_float=0) or tau_rand (as_float=1) of the
 * stream starting at state_host, i in [start, start+count), via GF(2) jump-ahead.          */
NND_API int nnd_rng_stream(nnd_ctx *ctx, const int64_t *state_host, int64_t start, int64_t count, int32_t as_float,
                   void *out, void *stream);
/* state after `n` draws (host-side jump-ahead) */
NND_API int nnd_rng_advance_host(nnd_ctx *ctx, int64_t *state_host, int64_t n);
/* pair distances in the canonical order: out[i] = dist(X[a[i]], X[b[i]]) -- the metric plugin evaluated outside a join
 * (distances.py named_distances / fast_distance_alternatives); used by the warm start and by the tests */
NND_API int nnd_pair_distances(nnd_ctx *ctx, const nnd_problem *P, const int32_t *a, const int32_t *b, int64_t n, float *out,
                       void *stream);
/* join-kernel profile, measured with CUDA events on the launching stream around every k_join
 * launch: out10 = {ms, launches, rows gathered, pairs evaluated, half-updates} for the local join
 * (utils.py:536-658) followed by the same five for the leaf join (pynndescent_.py:73-113).       */
NND_API int nnd_profile_read(nnd_ctx *ctx, double *out10, int32_t reset);
/* per-stage device time of the build path: with profiling on (nnd_stage_profile(ctx, 1), or NND_STAGE_PROFILE=1 in the
 * environment of nnd_create) every stage boundary records a CUDA event on the launching stream; nnd_stage_read waits for
 * the last one and returns the milliseconds spent in each stage since the last reset: out_ms[NND_STAGE_COUNT], indexed
 * by NND_ST_*.  (bench.py's stage table; the events cost about a microsecond of host time each.)                      */
#define NND_ST_OTHER 0        /* time between library calls, small bookkeeping kernels                                   */
#define NND_ST_FOREST_HYPER 1 /* rp_trees.py:323-420 hyperplanes of a level                                              */
#define NND_ST_FOREST_SIDE 2  /* margins / sides of a level (the pass that reads X)                                      */
#define NND_ST_FOREST_PART 3  /* scans, node bookkeeping, stable partition                                               */
#define NND_ST_LEAF_DIST 4    /* pynndescent_.py:73-113 leaf distance matrices (or the fused leaf kernel)                 */
#define NND_ST_LEAF_PUSH 5    /* pynndescent_.py:154-185 leaf pushes                                                     */
#define NND_ST_INIT_RANDOM 6
#define NND_ST_CAND_FILL 7    /* utils.py:261-306 rng tables, counts, scans, priorities + reverse-edge routing            */
#define NND_ST_CAND_SELECT 8  /* utils.py:289-318 heap selection, flag clearing                                          */
#define NND_ST_JOIN_PREP 9    /* batching prepass of the local join                                                      */
#define NND_ST_JOIN 10        /* utils.py:536-658 the local join kernel                                                  */
#define NND_ST_SCATTER 11     /* counting sort of the half-updates by target row                                         */
#define NND_ST_APPLY 12       /* utils.py:661-733 ordered replay of checked_flagged_heap_push                             */
#define NND_ST_DEHEAP 13      /* utils.py:189-218                                                                        */
#define NND_ST_EXCHANGE 14    /* sharded build: packing / unpacking of exchanged records and rows                         */
#define NND_STAGE_COUNT 16
NND_API int nnd_stage_profile(nnd_ctx *ctx, int32_t enable);
NND_API int nnd_stage_read(nnd_ctx *ctx, double *out_ms, int32_t reset);
/* EXPERIMENT (scripts/leaf_tc_experiment.py only; no product path calls it): the leaf join's pair matrix of leaves [l0, l1) as a
 * BF16 tensor-core Gram (mma.sync m16n8k16) used as a conservative lower bound of the squared distance.  counts_dev[3] (device,
 * accumulated): pairs, pairs the bound cannot exclude (lb < max(thr_i, thr_j)), and -- exact != 0 -- pairs whose exact distance
 * passes the reference's filter (pynndescent_.py:104).  Squared euclidean, d <= 128, leaves of at most 80 points.            */
NND_API int nnd_x_leaf_gram_tc(nnd_ctx *ctx, const nnd_problem *P, const int32_t *leaf_points, const int64_t *leaf_offsets, int64_t l0,
                       int64_t l1, int32_t max_leaf, int32_t exact, unsigned long long *counts_dev, void *stream);
/* kernels launched by this context since creation (bench.py's gpu_launches) */
NND_API int64_t nnd_launch_count(nnd_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* NND_B200_H */
