// k_query.cu -- seam 3: the query-time graph walk (pynndescent_.py:1770-1883 search_closure, with the tree
// descent :1716-1727 / rp_trees.py:2662-2677 select_side, utils.py:352-406 simple_heap_push, the visited set
// utils.py:323-349 and the heapq seed set) over PREPARED search structures in the reference's own format
// (CSR search graph, flat search tree, data in tree-leaf order).
//
// One warp per query.  Per-query state lives in shared memory: the query vector, the result max-heap (exact
// simple_heap_push layout), the seed min-heap ordered by (distance, index) like the reference's tuple heap, and
// an open-addressing visited set.  Distances of a leaf's / an expanded vertex's unvisited neighbours are
// evaluated lane-parallel (canonical sequential chain per pair), then APPLIED IN CSR ORDER with the bound
// tightened after every accepted push, exactly as :1859-1875 does.
//
// The reference threads ONE rng through all queries (:1796); it is only consumed on exact hyperplane ties and
// when a leaf is smaller than min(k, n_neighbors).  Pass 1 runs all queries in parallel and flags those that
// would need a draw (or overflow the on-chip state); pass 2 replays the flagged queries in query order on one
// warp with the real rng stream, global-memory seed heap and bitmap -- so results equal the reference's for
// every query, and the common case never serialises.
#include "nnd_common.cuh"

struct Seed { float d; int32_t i; };
__device__ __forceinline__ bool seed_lt(Seed a, Seed b) { return a.d < b.d || (a.d == b.d && a.i < b.i); }

struct QueryParams {
    const float *data; int64_t N; int d, ld; int metric; const void *aux;
    const int32_t *indptr, *indices;
    const float *hyper, *offsets; const int32_t *children, *tree_indices;
    const float *Q; int64_t nq; int k; double epsilon; float min_distance; int n_neighbors; int normalize;
    int32_t *out_idx; float *out_dist;
    uint8_t *flags;            // per query: 1 needs rng, 2 on-chip state overflow
    // pass 2 (sequential) resources
    const int32_t *todo; int64_t n_todo; long long *rng; Seed *g_seeds; int64_t g_seed_cap;
    uint32_t *g_bitmap; int32_t *g_touched; int64_t g_touched_cap;
    unsigned long long *stats; // [0] distance evaluations [1] pops
};

#define Q_WARPS 4
#define Q_SEEDCAP 768
#define Q_VISCAP 4096      // power of two; queries touching > half of it fall back to pass 2

__device__ __forceinline__ float warp_dot_canonical(const float *__restrict__ h, const float *__restrict__ x, int ld, int lane) {
    float acc = 0.0f;
    for (int c = lane * 4; c < ld; c += 128) {
        float4 a = *(const float4 *)(h + c), b = *(const float4 *)(x + c);
        acc = __fmaf_rn(a.x, b.x, acc); acc = __fmaf_rn(a.y, b.y, acc);
        acc = __fmaf_rn(a.z, b.z, acc); acc = __fmaf_rn(a.w, b.w, acc);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
    return acc;
}

// distance between the (normalised) query in shared memory and data row c, canonical sequential chain
__device__ __forceinline__ float query_distance(const QueryParams &P, const float *q, float qn2, double mq, double nq2, int32_t c) {
    const float *y = P.data + (int64_t)c * P.ld;
    if (P.metric == NND_METRIC_SQEUCLIDEAN) {
        float acc = 0.0f;
        for (int e = 0; e < P.d; ++e) { float df = __fsub_rn(q[e], y[e]); acc = __fmaf_rn(df, df, acc); }
        return acc;
    } else if (P.metric == NND_METRIC_CORRELATION) {
        const double *mu = (const double *)P.aux;
        const double my = mu[2 * (int64_t)c];
        double dp = 0.0;
        for (int e = 0; e < P.d; ++e) {
            double sx = __dadd_rn((double)q[e], -mq), sy = __dadd_rn((double)y[e], -my);
            dp = __fma_rn(sx, sy, dp);
        }
        return finish_correlation(dp, nq2, mu[2 * (int64_t)c + 1]);
    } else if (P.metric == NND_METRIC_ALT_COSINE || P.metric == NND_METRIC_ALT_DOT || P.metric == NND_METRIC_ALT_INNER) {
        float acc = 0.0f;
        for (int e = 0; e < P.d; ++e) acc = __fmaf_rn(q[e], y[e], acc);
        if (P.metric != NND_METRIC_ALT_COSINE) return finish_alt_dot(acc, P.metric);
        return finish_alt_cosine(acc, qn2, ((const float *)P.aux)[c]);
    }
    // SURVEY 8f-2 metrics: the same per-element steps and tails as the join (nnd_common.cuh); the query-side
    // auxiliaries arrive in qn2 (non-zero count / l1 norm)
    const int mc = nnd_metric_class(P.metric);
    if (mc == NND_METRIC_MINKOWSKI) {
        const double p = ((const double *)P.aux)[0], inv_p = ((const double *)P.aux)[1];
        double acc = 0.0;
        for (int e = 0; e < P.d; ++e) acc = nnd_minkowski_step(acc, q[e], y[e], p);
        return finish_minkowski(acc, inv_p);
    }
    if (mc == NND_METRIC_MANHATTAN) {
        double acc = 0.0;
        for (int e = 0; e < P.d; ++e) acc = nnd_metric_step<NND_METRIC_MANHATTAN, double>(acc, q[e], y[e]);
        return finish_plain(P.metric, acc, P.d);
    } else if (mc == NND_METRIC_CHEBYSHEV) {
        float acc = 0.0f;
        for (int e = 0; e < P.d; ++e) acc = nnd_metric_step<NND_METRIC_CHEBYSHEV, float>(acc, q[e], y[e]);
        return acc;
    } else if (mc == NND_METRIC_CANBERRA) {
        double acc = 0.0;
        for (int e = 0; e < P.d; ++e) acc = nnd_metric_step<NND_METRIC_CANBERRA, double>(acc, q[e], y[e]);
        return finish_plain(P.metric, acc, P.d);
    } else if (mc == NND_METRIC_HAMMING) {
        float acc = 0.0f;
        for (int e = 0; e < P.d; ++e) acc = nnd_metric_step<NND_METRIC_HAMMING, float>(acc, q[e], y[e]);
        return finish_plain(P.metric, (double)acc, P.d);
    } else if (mc == NND_METRIC_ALT_JACCARD) {
        float acc = 0.0f;
        for (int e = 0; e < P.d; ++e) acc = nnd_metric_step<NND_METRIC_ALT_JACCARD, float>(acc, q[e], y[e]);
        return finish_binary(P.metric, acc, qn2, ((const float *)P.aux)[c], P.d);
    } else {
        float acc = 0.0f;
        for (int e = 0; e < P.d; ++e) acc = nnd_metric_step<NND_METRIC_ALT_HELLINGER, float>(acc, q[e], y[e]);
        return finish_alt_hellinger(acc, qn2, ((const float *)P.aux)[c]);
    }
}

// utils.py:352-406 simple_heap_push (no duplicate test), executed by lane 0
__device__ __forceinline__ void simple_push_lane0(float *pri, int32_t *ind, int size, float p, int32_t n) {
    if (p >= pri[0]) return;
    int i = 0, i_swap;
    for (;;) {
        int ic1 = 2 * i + 1, ic2 = ic1 + 1;
        if (ic1 >= size) break;
        else if (ic2 >= size) { if (pri[ic1] > p) i_swap = ic1; else break; }
        else if (pri[ic1] >= pri[ic2]) { if (p < pri[ic1]) i_swap = ic1; else break; }
        else { if (p < pri[ic2]) i_swap = ic2; else break; }
        pri[i] = pri[i_swap]; ind[i] = ind[i_swap];
        i = i_swap;
    }
    pri[i] = p; ind[i] = n;
}
// CPython heapq.heappush / heappop on (d, i) tuples, lane 0
__device__ __forceinline__ void seed_push(Seed *h, int &n, Seed x) {
    int pos = n++;
    while (pos > 0) { int parent = (pos - 1) >> 1; if (seed_lt(x, h[parent])) { h[pos] = h[parent]; pos = parent; } else break; }
    h[pos] = x;
}
__device__ __forceinline__ Seed seed_pop(Seed *h, int &n) {
    Seed last = h[--n];
    if (n == 0) return last;
    Seed ret = h[0];
    int pos = 0, child = 1;
    while (child < n) {
        int right = child + 1;
        if (right < n && !seed_lt(h[child], h[right])) child = right;
        h[pos] = h[child]; pos = child; child = 2 * pos + 1;
    }
    while (pos > 0) { int parent = (pos - 1) >> 1; if (seed_lt(last, h[parent])) { h[pos] = h[parent]; pos = parent; } else break; }
    h[pos] = last;
    return ret;
}

// visited set: returns true if c was ALREADY visited; marks it otherwise (utils.py:338-349)
__device__ __forceinline__ bool visit(int32_t c, int32_t *vis, uint32_t *bitmap, int32_t *touched, int *n_touched, int64_t touched_cap,
                                      bool *overflow) {
    if (bitmap) {
        uint32_t old = atomicOr(&bitmap[c >> 5], 1u << (c & 31));
        if (old & (1u << (c & 31))) return true;
        int t = atomicAdd(n_touched, 1);
        if (t < touched_cap) touched[t] = c; else *overflow = true;
        return false;
    }
    uint32_t h = ((uint32_t)c * 2654435761u) >> (32 - 12);   // log2(Q_VISCAP) = 12
    for (int probe = 0; probe < Q_VISCAP; ++probe) {
        int32_t old = atomicCAS(&vis[h], -1, c);
        if (old == -1) { atomicAdd(n_touched, 1); return false; }
        if (old == c) return true;
        h = (h + 1) & (Q_VISCAP - 1);
    }
    *overflow = true;
    return true;
}

// one query by one warp.  Returns 0 ok, 1 needs the rng (only when rng == nullptr), 2 state overflow.
__device__ int run_query(const QueryParams &P, int64_t qi, float *q, float *pri, int32_t *ind, Seed *seeds, int seed_cap,
                         int32_t *vis, int *s_cnt, long long *rng, int lane) {
    const int k = P.k, ld = P.ld;
    uint32_t *bitmap = rng ? P.g_bitmap : nullptr;
    int32_t *touched = rng ? P.g_touched : nullptr;
    for (int j = lane; j < k; j += 32) { pri[j] = CUDART_INF_F; ind[j] = -1; }
    if (!bitmap) for (int j = lane; j < Q_VISCAP; j += 32) vis[j] = -1;
    if (lane == 0) { s_cnt[0] = 0; s_cnt[1] = 0; }   // [0] visited count [1] overflow flag
    // ---- query vector (normalised for cosine/dot, :1806-1811)
    const float *qq = P.Q + qi * ld;
    float inv_ok = 1.0f;
    if (P.normalize) {
        float s = 0.0f;   // sequential f32 sum of squares (every lane computes the same value)
        for (int e = 0; e < P.d; ++e) s = __fadd_rn(s, __fmul_rn(qq[e], qq[e]));
        float norm = __fsqrt_rn(s);
        if (!(norm > 0.0f)) inv_ok = 0.0f;
        for (int e = lane; e < ld; e += 32) q[e] = (e < P.d && inv_ok != 0.0f) ? __fdiv_rn(qq[e], norm) : 0.0f;
    } else {
        for (int e = lane; e < ld; e += 32) q[e] = qq[e];
    }
    __syncwarp();
    if (inv_ok == 0.0f) return 0;   // zero query under cosine/dot: the reference `continue`s (row stays -1 / inf)
    float qn2 = 0.0f; double mq = 0.0, nq2 = 0.0;
    if (P.metric == NND_METRIC_ALT_COSINE) for (int e = 0; e < P.d; ++e) qn2 = __fmaf_rn(q[e], q[e], qn2);
    if (nnd_metric_class(P.metric) == NND_METRIC_ALT_JACCARD) for (int e = 0; e < P.d; ++e) if (q[e] != 0.0f) qn2 += 1.0f;
    if (P.metric == NND_METRIC_ALT_HELLINGER) for (int e = 0; e < P.d; ++e) qn2 = __fadd_rn(qn2, q[e]);
    if (P.metric == NND_METRIC_CORRELATION) {
        double sq = 0.0;
        for (int e = 0; e < P.d; ++e) sq = __dadd_rn(sq, (double)q[e]);
        mq = __ddiv_rn(sq, (double)P.d);
        for (int e = 0; e < P.d; ++e) { double sx = __dadd_rn((double)q[e], -mq); nq2 = __fma_rn(sx, sx, nq2); }
    }
    bool ovf = false;
    unsigned long long n_dist = 0, n_pop = 0;
    // ---- tree descent (:1716-1727)
    int32_t node = 0;
    while (P.children[2 * node] > 0) {
        float margin = __fadd_rn(P.offsets[node], warp_dot_canonical(P.hyper + (int64_t)node * ld, q, ld, lane));
        int side;
        if (fabs((double)margin) < 1e-8) {
            if (!rng) return 1;
            int32_t r = 0;
            if (lane == 0) r = tau_step(rng);
            r = __shfl_sync(0xffffffffu, r, 0);
            long long a = (r == (int32_t)0x80000000) ? (long long)r : (long long)(r < 0 ? -r : r);
            side = (int)((a % 2) != 0);
        } else side = margin > 0 ? 0 : 1;
        node = P.children[2 * node + side];
    }
    const int32_t b0 = -P.children[2 * node], b1 = -P.children[2 * node + 1];
    const int n_random = min(k, P.n_neighbors) - (b1 - b0);
    if (n_random > 0 && !rng) return 1;
    int ns = 0;
    // ---- seed from the leaf (:1826-1832): lane-parallel distances, applied in order
    for (int32_t c0 = b0; c0 < b1; c0 += 32) {
        int32_t c = -1; float dd = 0.f;
        if (c0 + lane < b1) { c = P.tree_indices[c0 + lane]; dd = query_distance(P, q, qn2, mq, nq2, c); ++n_dist; visit(c, vis, bitmap, touched, s_cnt, P.g_touched_cap, &ovf); }
        int cnt = min(32, b1 - c0);
        for (int s = 0; s < cnt; ++s) {
            float dv = __shfl_sync(0xffffffffu, dd, s); int32_t cv = __shfl_sync(0xffffffffu, c, s);
            if (lane == 0) {
                simple_push_lane0(pri, ind, k, dv, cv);
                if (ns < seed_cap) { Seed x; x.d = dv; x.i = cv; seed_push(seeds, ns, x); } else s_cnt[1] = 1;
            }
        }
        __syncwarp();
    }
    // ---- random seeds when the leaf is short (:1835-1847), sequential rng
    for (int j = 0; j < n_random; ++j) {
        int32_t r = 0;
        if (lane == 0) r = tau_step(rng);
        r = __shfl_sync(0xffffffffu, r, 0);
        long long a = (r == (int32_t)0x80000000) ? (long long)r : (long long)(r < 0 ? -r : r);
        long long mm = a % P.N; if (mm < 0) mm += P.N;
        int32_t c = (int32_t)mm;
        bool was = false;
        if (lane == 0) was = visit(c, vis, bitmap, touched, s_cnt, P.g_touched_cap, &ovf);
        was = __shfl_sync(0xffffffffu, (int)was, 0);
        if (!was) {
            float dd = 0.f;
            if (lane == 0) {
                dd = query_distance(P, q, qn2, mq, nq2, c); ++n_dist;
                simple_push_lane0(pri, ind, k, dd, c);
                if (ns < seed_cap) { Seed x; x.d = dd; x.i = c; seed_push(seeds, ns, x); } else s_cnt[1] = 1;
            }
        }
        __syncwarp();
    }
    // ---- best-first walk (:1849-1881)
    // :1850-1852 in float64 then stored to the float32 local (two roundings, as numba does it)
    auto bound_of = [&](float top) { return (float)__dadd_rn((double)top, __dmul_rn(P.epsilon, __dadd_rn((double)top, -(double)P.min_distance))); };
    float bound = bound_of(pri[0]);
    ns = __shfl_sync(0xffffffffu, ns, 0);
    if (ns == 0) return 0;
    Seed cur; cur.d = 0.f; cur.i = 0;
    if (lane == 0) cur = seed_pop(seeds, ns);
    cur.d = __shfl_sync(0xffffffffu, cur.d, 0); cur.i = __shfl_sync(0xffffffffu, cur.i, 0);
    ++n_pop;
    while (cur.d < bound) {
        const int32_t e0 = P.indptr[cur.i], e1 = P.indptr[cur.i + 1];
        for (int32_t j0 = e0; j0 < e1; j0 += 32) {
            int32_t c = -1; float dd = 0.f; bool fresh = false;
            if (j0 + lane < e1) {
                c = P.indices[j0 + lane];
                fresh = !visit(c, vis, bitmap, touched, s_cnt, P.g_touched_cap, &ovf);
                if (fresh) { dd = query_distance(P, q, qn2, mq, nq2, c); ++n_dist; }
            }
            unsigned mk = __ballot_sync(0xffffffffu, fresh);
            while (mk) {                                   // apply in CSR order, the bound tightens as we go
                int s = __ffs(mk) - 1; mk &= mk - 1;
                float dv = __shfl_sync(0xffffffffu, dd, s); int32_t cv = __shfl_sync(0xffffffffu, c, s);
                if (dv < bound) {
                    if (lane == 0) {
                        simple_push_lane0(pri, ind, k, dv, cv);
                        if (ns < seed_cap) { Seed x; x.d = dv; x.i = cv; seed_push(seeds, ns, x); } else s_cnt[1] = 1;
                    }
                    __syncwarp();
                    bound = bound_of(pri[0]);
                }
            }
        }
        ns = __shfl_sync(0xffffffffu, ns, 0);
        if (ns == 0) break;
        if (lane == 0) cur = seed_pop(seeds, ns);
        cur.d = __shfl_sync(0xffffffffu, cur.d, 0); cur.i = __shfl_sync(0xffffffffu, cur.i, 0);
        ++n_pop;
        if (!bitmap && (s_cnt[0] > Q_VISCAP / 2)) ovf = true;
        if (__any_sync(0xffffffffu, ovf) || s_cnt[1]) return 2;
    }
    if (__any_sync(0xffffffffu, ovf) || s_cnt[1]) return 2;
    if (P.stats) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) n_dist += __shfl_xor_sync(0xffffffffu, n_dist, o);
        if (lane == 0) { atomicAdd(&P.stats[0], n_dist); atomicAdd(&P.stats[1], n_pop); }
    }
    return 0;
}

__global__ void __launch_bounds__(Q_WARPS * 32) k_query(QueryParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t per = ((size_t)P.ld * 4 + (size_t)P.k * 8 + (size_t)Q_SEEDCAP * 8 + (size_t)Q_VISCAP * 4 + 16 + 15) & ~size_t(15);
    unsigned char *base = smem_raw + (size_t)warp * per;
    float *q = (float *)base; float *pri = q + P.ld; int32_t *ind = (int32_t *)(pri + P.k);
    Seed *seeds = (Seed *)(ind + P.k);   // offset ld*4 + 8k: 8-byte aligned
    int32_t *vis = (int32_t *)(seeds + Q_SEEDCAP); int *s_cnt = (int *)(vis + Q_VISCAP);
    const int64_t qi = (int64_t)blockIdx.x * Q_WARPS + warp;
    if (qi >= P.nq) return;
    int st = run_query(P, qi, q, pri, ind, seeds, Q_SEEDCAP, vis, s_cnt, nullptr, lane);
    __syncwarp();
    if (lane == 0) P.flags[qi] = (uint8_t)st;
    if (st == 0)
        for (int j = lane; j < P.k; j += 32) { P.out_idx[qi * P.k + j] = ind[j]; P.out_dist[qi * P.k + j] = pri[j]; }
}

// pass 2: flagged queries, ascending query order, one warp, real rng, global seed heap + bitmap
__global__ void __launch_bounds__(32) k_query_seq(QueryParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    float *q = (float *)smem_raw; float *pri = q + P.ld; int32_t *ind = (int32_t *)(pri + P.k);
    int *s_cnt = (int *)(ind + P.k);
    for (int64_t t = 0; t < P.n_todo; ++t) {
        const int64_t qi = P.todo[t];
        int st = run_query(P, qi, q, pri, ind, P.g_seeds, (int)min((int64_t)0x7fffffff, P.g_seed_cap), nullptr, s_cnt, P.rng, lane);
        __syncwarp();
        // clear the bits this query set
        int nt = s_cnt[0];
        for (int j = lane; j < nt && j < P.g_touched_cap; j += 32) { int32_t c = P.g_touched[j]; atomicAnd(&P.g_bitmap[c >> 5], ~(1u << (c & 31))); }
        __syncwarp();
        if (lane == 0) P.flags[qi] = (uint8_t)(st == 0 ? 0 : 4);
        for (int j = lane; j < P.k; j += 32) { P.out_idx[qi * P.k + j] = ind[j]; P.out_dist[qi * P.k + j] = pri[j]; }
        __syncwarp();
    }
}

__global__ void k_query_collect(const uint8_t *flags, int64_t nq, int32_t *todo, unsigned long long *n_todo) {
    // single thread: keeps ascending query order (the list is tiny in practice)
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long n = 0;
        for (int64_t i = 0; i < nq; ++i) if (flags[i]) todo[n++] = (int32_t)i;
        *n_todo = n;
    }
}

extern "C" int nnd_query_workspace_bytes(int64_t N, int64_t nq, int32_t k, size_t *bytes) {
    NND_CHECK(bytes && N > 0 && nq >= 0 && k > 0, NND_ERR_INVALID, "nnd_query_workspace_bytes: bad args");
    WsAlloc A(nullptr, 0);
    A.get<uint8_t>(nq + 16); A.get<int32_t>(nq + 16); A.get<unsigned long long>(8); A.get<long long>(4);
    A.get<Seed>(N + 64); A.get<uint32_t>(N / 32 + 64); A.get<int32_t>(N + 64);
    *bytes = A.off + 4096;
    return NND_OK;
}

extern "C" int nnd_query(nnd_ctx *ctx, const nnd_search_index *S, const float *Q, int64_t nq, int32_t k, double epsilon,
                         int64_t *rng_state_host, int32_t *out_idx, float *out_dist, void *ws, size_t ws_bytes,
                         int64_t *stats_host, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    NND_CHECK(ctx && S && Q && out_idx && out_dist && rng_state_host && ws && nq >= 0 && k > 0, NND_ERR_INVALID, "nnd_query: bad args");
    NND_CHECK(S->metric >= 0 && S->metric <= NND_METRIC_MAX, NND_ERR_UNSUPPORTED, "nnd_query: metric id %d", S->metric);
    NND_CHECK(S->ld % 4 == 0 && S->ld >= S->d, NND_ERR_INVALID, "nnd_query: ld must be a multiple of 4 and >= d");
    if (nq == 0) return NND_OK;
    WsAlloc A(ws, ws_bytes);
    uint8_t *flags = A.get<uint8_t>(nq + 16); int32_t *todo = A.get<int32_t>(nq + 16);
    unsigned long long *scal = A.get<unsigned long long>(8); long long *d_rng = A.get<long long>(4);
    Seed *g_seeds = A.get<Seed>(S->N + 64); uint32_t *g_bitmap = A.get<uint32_t>(S->N / 32 + 64); int32_t *g_touched = A.get<int32_t>(S->N + 64);
    NND_CHECK(A.ok, NND_ERR_WORKSPACE, "nnd_query: workspace too small");
    QueryParams P; memset(&P, 0, sizeof(P));
    P.data = S->data; P.N = S->N; P.d = S->d; P.ld = S->ld; P.metric = S->metric; P.aux = S->aux;
    P.indptr = S->indptr; P.indices = S->indices; P.hyper = S->hyperplanes; P.offsets = S->offsets; P.children = S->children;
    P.tree_indices = S->tree_indices; P.Q = Q; P.nq = nq; P.k = k; P.epsilon = epsilon; P.min_distance = S->min_distance;
    P.n_neighbors = S->n_neighbors; P.normalize = S->normalize; P.out_idx = out_idx; P.out_dist = out_dist; P.flags = flags;
    P.stats = scal + 1;
    NND_CUDA(cudaMemsetAsync(scal, 0, sizeof(unsigned long long) * 8, st));
    const size_t per = ((size_t)P.ld * 4 + (size_t)k * 8 + (size_t)Q_SEEDCAP * 8 + (size_t)Q_VISCAP * 4 + 16 + 15) & ~size_t(15);
    const size_t smem = per * Q_WARPS + 16;
    NND_CHECK(smem <= 200 * 1024, NND_ERR_UNSUPPORTED, "nnd_query: d=%d k=%d need %zu B of shared memory", S->d, k, smem);
    NND_CUDA(cudaFuncSetAttribute(k_query, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_query<<<(unsigned)cdiv(nq, Q_WARPS), Q_WARPS * 32, smem, st>>>(P); NND_LAUNCH(ctx);
    k_query_collect<<<1, 32, 0, st>>>(flags, nq, todo, scal); NND_LAUNCH(ctx);
    unsigned long long n_todo = 0;
    NND_CUDA(cudaMemcpyAsync(&n_todo, scal, sizeof(n_todo), cudaMemcpyDeviceToHost, st));
    NND_CUDA(cudaStreamSynchronize(st));
    // the rng copy of this call (:1796): consumed only by the flagged queries, in query order
    long long h_rng[3] = {rng_state_host[0], rng_state_host[1], rng_state_host[2]};
    if (n_todo > 0) {
        NND_CUDA(cudaMemcpyAsync(d_rng, h_rng, sizeof(h_rng), cudaMemcpyHostToDevice, st));
        NND_CUDA(cudaMemsetAsync(g_bitmap, 0, sizeof(uint32_t) * (S->N / 32 + 64), st));
        P.todo = todo; P.n_todo = (int64_t)n_todo; P.rng = d_rng; P.g_seeds = g_seeds; P.g_seed_cap = S->N + 64;
        P.g_bitmap = g_bitmap; P.g_touched = g_touched; P.g_touched_cap = S->N + 64;
        const size_t smem2 = (size_t)P.ld * 4 + (size_t)k * 8 + 64;
        NND_CUDA(cudaFuncSetAttribute(k_query_seq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        k_query_seq<<<1, 32, smem2, st>>>(P); NND_LAUNCH(ctx);
        NND_CUDA(cudaStreamSynchronize(st));
    }
    unsigned long long hs[3];
    NND_CUDA(cudaMemcpyAsync(hs, scal, sizeof(hs), cudaMemcpyDeviceToHost, st));
    NND_CUDA(cudaStreamSynchronize(st));
    if (stats_host) { stats_host[0] = (int64_t)hs[1]; stats_host[1] = (int64_t)hs[2]; stats_host[2] = (int64_t)n_todo; }
    NND_CUDA(cudaGetLastError());
    return NND_OK;
}
