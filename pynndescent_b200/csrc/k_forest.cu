// k_forest.cu -- level-synchronous random-projection forest (seam 1).
//
// Reference: rp_trees.py:2815-2888 make_forest -> :2515-2554 make_dense_tree -> :2177-2302
// make_euclidean_tree / make_angular_tree -> :323-420 / :60-171 *_random_projection_split, and
// :2891-2922 rptree_leaf_array.
//
// Same split rule (two random members define the hyperplane; margin>0 -> left; |margin|<1e-8 ->
// coin; all-on-one-side -> random split) and the same STABLE partition as the reference, so a
// leaf's points stay in ascending id order and leaves come out in DFS left-to-right order.  The
// reference draws its random numbers from ONE sequential stream per tree in DFS order, which
// serialises the whole tree; here every node draws from hash(tree seed, level, ordinal among the
// tree's active nodes of that level) so that all nodes of a level -- of all trees -- split in one
// pass over the points (the formulation of the reference's own threaded_rp_trees.py:12-44
// prototype).  oracle/nnd_oracle.c level_tree() restates this scheme; the CUDA result is
// bit-identical to it.  Trees are statistically equivalent to the reference's, not identical.
#include <vector>

#include "nnd_common.cuh"

struct ForestWs {
    int32_t *perm[2];     // [TN] point id at each position (ping-pong)
    int32_t *seg[2];      // [TN] active-node ordinal of the position, -1 if already a final leaf
    uint8_t *side;        // [TN]
    uint8_t *segstart;    // [TN] 1 where a segment starts
    uint32_t *flag32;     // [TN+1] scan input
    uint32_t *pl;         // [TN+2] scan output
    uint32_t *scan_tmp;
    // node tables (ping-pong): start, size
    int32_t *nstart[2], *nsize[2];
    int32_t *nleft;       // [M]
    uint8_t *ndegen;      // [M]
    uint32_t *nchild;     // [M+1] number of active children (0..2) -> scan
    uint32_t *nchild_scan;
    float *hyper;         // [M][ld]
    float *noff;          // [M]
    int32_t *tree_first;  // [n_trees]
    uint64_t *tree_seed;  // [n_trees]
    int32_t *scalars;     // [0] n_degenerate [1] max_leaf [2] next active count
    int32_t *act[2];      // [TN] active positions (ping-pong, unordered)
    int32_t *invp;        // [TN] position of point p in tree t at invp[t*N+p], -1 once it sits in a final leaf (NND_FOREST_PM)
    int64_t M_max;
};

static int carve_forest_ws(void *ws, size_t ws_bytes, int64_t N, int ld, int n_trees, ForestWs *W, size_t *need) {
    WsAlloc A(ws, ws_bytes);
    int64_t TN = (int64_t)n_trees * N;
    W->M_max = TN / 2 + n_trees + 16;  // a node splits only if size > leaf_size >= 1 ... at least 2 points
    for (int i = 0; i < 2; ++i) { W->perm[i] = A.get<int32_t>(TN); W->seg[i] = A.get<int32_t>(TN); }
    W->side = A.get<uint8_t>(TN); W->segstart = A.get<uint8_t>(TN + 1);
    W->flag32 = A.get<uint32_t>(TN + 1); W->pl = A.get<uint32_t>(TN + 2);
    W->scan_tmp = A.get<uint32_t>(nnd_scan_tmp_elems(TN + 2));
    for (int i = 0; i < 2; ++i) { W->nstart[i] = A.get<int32_t>(W->M_max); W->nsize[i] = A.get<int32_t>(W->M_max); }
    W->nleft = A.get<int32_t>(W->M_max); W->ndegen = A.get<uint8_t>(W->M_max);
    W->nchild = A.get<uint32_t>(W->M_max + 1); W->nchild_scan = A.get<uint32_t>(W->M_max + 2);
    W->noff = A.get<float>(W->M_max);
    W->tree_first = A.get<int32_t>(n_trees); W->tree_seed = A.get<uint64_t>(n_trees);
    W->scalars = A.get<int32_t>(16);
    for (int i = 0; i < 2; ++i) W->act[i] = A.get<int32_t>(TN);
    W->invp = A.get<int32_t>(TN);
    // hyperplanes: nodes that split hold > leaf_size points, but leaf_size is not known here;
    // active nodes of one level are disjoint and each has >= 2 points
    W->hyper = A.get<float>((size_t)0);  // sized by forest_hyper_elems() below (needs leaf_size)
    if (need) *need = A.off;
    (void)ld;
    return A.ok ? NND_OK : NND_ERR_WORKSPACE;
}

extern "C" int nnd_forest_workspace_bytes(int64_t N, int32_t d, int32_t n_trees, size_t *bytes) {
    NND_CHECK(bytes && N > 0 && n_trees > 0, NND_ERR_INVALID, "nnd_forest_workspace_bytes: bad args");
    ForestWs W; size_t need = 0;
    int ld = (d + 3) & ~3;
    carve_forest_ws(nullptr, 0, N, ld, n_trees, &W, &need);
    // hyperplane pool: active nodes of a level hold >= 61 points each in every reference
    // configuration (leaf_size >= 60, rp_trees.py:2845-2846); sized for leaf_size >= 8 generally
    size_t hyper = ((size_t)n_trees * N / 8 + n_trees + 16) * (size_t)ld * 4;
    *bytes = need + hyper + 8192;
    return NND_OK;
}

// ---------------------------------------------------------------- canonical warp reductions
__device__ __forceinline__ float warp_lane_dot(const float *__restrict__ h, const float *__restrict__ x, int ld, int lane) {
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
// utils.py:75-90 norm(): f32 products, f64 accumulation, f64 sqrt, f4 result
__device__ __forceinline__ float warp_lane_norm(const float *__restrict__ x, int ld, int lane) {
    double acc = 0.0;
    for (int c = lane * 4; c < ld; c += 128) {
        float4 a = *(const float4 *)(x + c);
        acc = __dadd_rn(acc, (double)__fmul_rn(a.x, a.x)); acc = __dadd_rn(acc, (double)__fmul_rn(a.y, a.y));
        acc = __dadd_rn(acc, (double)__fmul_rn(a.z, a.z)); acc = __dadd_rn(acc, (double)__fmul_rn(a.w, a.w));
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
    return (float)__dsqrt_rn(acc);
}

__device__ __forceinline__ uint64_t node_key(uint64_t seed, int level, int64_t ordinal) {
    return mix64(seed ^ mix64(((uint64_t)level << 40) ^ (uint64_t)ordinal));
}

// ---------------------------------------------------------------- kernels
__global__ void k_forest_init(int32_t *perm, int32_t *seg, uint8_t *segstart, int32_t *act, int32_t *leaf_points,
                              uint32_t *flag32, int64_t N, int n_trees, int active, int32_t *invp) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_trees * N) return;
    int64_t t = i / N, p = i - t * N;
    perm[i] = (int32_t)p;
    seg[i] = active ? (int32_t)t : -1;
    segstart[i] = (p == 0);
    act[i] = (int32_t)i;
    flag32[i] = 0;
    if (!active) leaf_points[i] = (int32_t)p;
    if (invp) invp[i] = active ? (int32_t)i : -1;
}

// one warp per active node: choose the two members, build the hyperplane (rp_trees.py:351-367 / :88-118)
__global__ void k_forest_hyper(const float *X, int ld, int64_t N, int angular, int level, int M, const int32_t *nstart,
                               const int32_t *nsize, const int32_t *perm, const int32_t *tree_first,
                               const uint64_t *tree_seed, float *hyper, float *noff) {
    extern __shared__ float s_t[];  // [warps][ld]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int o = blockIdx.x * (blockDim.x >> 5) + warp;
    if (o >= M) return;
    const int start = nstart[o], n = nsize[o];
    const int tree = (int)(start / N);
    const uint64_t key = node_key(tree_seed[tree], level, (int64_t)(o - tree_first[tree]));
    int64_t li = (int64_t)(mix64(key ^ 0xAULL) % (uint64_t)n);
    int64_t ri = (int64_t)(mix64(key ^ 0xBULL) % (uint64_t)n);
    ri += (li == ri);
    ri = ri % n;
    const float *xl = X + (int64_t)perm[start + li] * ld, *xr = X + (int64_t)perm[start + ri] * ld;
    float *h = hyper + (size_t)o * ld;
    float *t = s_t + (size_t)warp * ld;
    if (!angular) {
        for (int e = lane; e < ld; e += 32) {
            float he = __fsub_rn(xl[e], xr[e]);
            h[e] = he;
            t[e] = __fmul_rn(he, __fadd_rn(xl[e], xr[e]));
        }
        __syncwarp();
        if (lane == 0) {
            float ofs = 0.0f;
            for (int e = 0; e < ld; ++e) ofs = __fsub_rn(ofs, __fmul_rn(0.5f, t[e]));
            noff[o] = ofs;
        }
    } else {
        float ln = warp_lane_norm(xl, ld, lane), rn = warp_lane_norm(xr, ld, lane);
        if (fabs((double)ln) < 1e-8) ln = 1.0f;
        if (fabs((double)rn) < 1e-8) rn = 1.0f;
        for (int e = lane; e < ld; e += 32) t[e] = __fsub_rn(__fdiv_rn(xl[e], ln), __fdiv_rn(xr[e], rn));
        __syncwarp();
        float hn = warp_lane_norm(t, ld, lane);
        if (fabs((double)hn) < 1e-8) hn = 1.0f;
        for (int e = lane; e < ld; e += 32) h[e] = __fdiv_rn(t[e], hn);
        if (lane == 0) noff[o] = 0.0f;
    }
}

// one warp per SIDE_PW active positions: side of the hyperplane (rp_trees.py:375-391)
#define SIDE_PW 2
__global__ void __launch_bounds__(256, 6) k_forest_side(const float *X, int ld, int64_t n_act, const int32_t *act, int64_t N, int level,
                              const int32_t *perm, const int32_t *seg, const int32_t *nstart, const int32_t *tree_first,
                              const uint64_t *tree_seed, const float *hyper, const float *noff, uint8_t *side,
                              uint32_t *flag32) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t i0 = w * SIDE_PW;
    if (i0 >= n_act) return;
    int32_t pos[SIDE_PW], o[SIDE_PW];
    const float *x[SIDE_PW];
#pragma unroll
    for (int q = 0; q < SIDE_PW; ++q) {
        int64_t i = i0 + q < n_act ? i0 + q : n_act - 1;
        pos[q] = act[i]; o[q] = seg[pos[q]];
        x[q] = X + (int64_t)perm[pos[q]] * ld;
    }
    float acc[SIDE_PW];
#pragma unroll
    for (int q = 0; q < SIDE_PW; ++q) acc[q] = 0.0f;
    for (int c = lane * 4; c < ld; c += 128) {
        float4 b[SIDE_PW], a[SIDE_PW];
#pragma unroll
        for (int q = 0; q < SIDE_PW; ++q) { b[q] = *(const float4 *)(x[q] + c); a[q] = *(const float4 *)(hyper + (size_t)o[q] * ld + c); }
#pragma unroll
        for (int q = 0; q < SIDE_PW; ++q) {
            acc[q] = __fmaf_rn(a[q].x, b[q].x, acc[q]); acc[q] = __fmaf_rn(a[q].y, b[q].y, acc[q]);
            acc[q] = __fmaf_rn(a[q].z, b[q].z, acc[q]); acc[q] = __fmaf_rn(a[q].w, b[q].w, acc[q]);
        }
    }
#pragma unroll
    for (int q = 0; q < SIDE_PW; ++q) {
#pragma unroll
        for (int sft = 16; sft >= 1; sft >>= 1) acc[q] = __fadd_rn(acc[q], __shfl_xor_sync(0xffffffffu, acc[q], sft));
    }
    if (lane < SIDE_PW && i0 + lane < n_act) {
        float accq = acc[0]; int32_t pq = pos[0], oq = o[0];
#pragma unroll
        for (int q = 1; q < SIDE_PW; ++q) if (lane == q) { accq = acc[q]; pq = pos[q]; oq = o[q]; }
        float margin = __fadd_rn(noff[oq], accq);
        uint8_t s;
        if (fabs((double)margin) < 1e-8) {
            const int start = nstart[oq];
            const int tree = (int)(start / N);
            const uint64_t key = node_key(tree_seed[tree], level, (int64_t)(oq - tree_first[tree]));
            s = (uint8_t)(mix64(key ^ (0x100ULL + (uint64_t)(pq - start))) & 1ULL);
        } else s = margin > 0 ? 0 : 1;
        side[pq] = s;
        flag32[pq] = (s == 0);
    }
}


// ld <= 128 fast path: one warp walks SIDE_RUN consecutive active entries; the node's hyperplane stays in
// registers (one float4 per lane) while consecutive entries belong to the same node, which halves the L2
// traffic of the pass (x is the only per-position stream left).
#define SIDE_RUN 16
__global__ void __launch_bounds__(256) k_forest_side128(const float *X, int ld, int64_t n_act, const int32_t *act, int64_t N,
                                                        int level, const int32_t *perm, const int32_t *seg,
                                                        const int32_t *nstart, const int32_t *tree_first,
                                                        const uint64_t *tree_seed, const float *hyper, const float *noff,
                                                        uint8_t *side, uint32_t *flag32) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t i0 = w * SIDE_RUN;
    if (i0 >= n_act) return;
    const int cnt = (int)min((int64_t)SIDE_RUN, n_act - i0);
    const bool has = lane * 4 < ld;
    // lanes 0..cnt-1 fetch the per-entry scalars, then broadcast
    int32_t my_pos = 0, my_o = 0, my_pt = 0;
    if (lane < cnt) { my_pos = act[i0 + lane]; my_o = seg[my_pos]; my_pt = perm[my_pos]; }
    int cur_o = -1; float4 h4 = make_float4(0.f, 0.f, 0.f, 0.f); float off = 0.f;
    // next row prefetched one entry ahead; two NAMED registers (a dynamically indexed float4[2] lives in local memory: one STL
    // and one LDL per entry on top of the row itself)
    float4 xcur, xnext;
    {
        int pt0 = __shfl_sync(0xffffffffu, my_pt, 0);
        xcur = has ? *(const float4 *)(X + (int64_t)pt0 * ld + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    xnext = xcur;
    uint32_t sbits = 0;   // lane 0 collects the side bits of the run
    for (int q = 0; q < cnt; ++q) {
        if (q + 1 < cnt) {
            int ptn = __shfl_sync(0xffffffffu, my_pt, q + 1);
            xnext = has ? *(const float4 *)(X + (int64_t)ptn * ld + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const int o = __shfl_sync(0xffffffffu, my_o, q);
        if (o != cur_o) {
            cur_o = o;
            h4 = has ? *(const float4 *)(hyper + (size_t)o * ld + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            off = noff[o];
        }
        const float4 b = xcur;
        xcur = xnext;
        float acc = 0.0f;
        acc = __fmaf_rn(h4.x, b.x, acc); acc = __fmaf_rn(h4.y, b.y, acc); acc = __fmaf_rn(h4.z, b.z, acc); acc = __fmaf_rn(h4.w, b.w, acc);
#pragma unroll
        for (int sft = 16; sft >= 1; sft >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, sft));
        const float margin = __fadd_rn(off, acc);
        uint32_t s;
        if (fabs((double)margin) < 1e-8) {
            const int pos = __shfl_sync(0xffffffffu, my_pos, q);
            const int start = nstart[o];
            const int tree = (int)(start / N);
            const uint64_t key = node_key(tree_seed[tree], level, (int64_t)(o - tree_first[tree]));
            s = (uint32_t)(mix64(key ^ (0x100ULL + (uint64_t)(pos - start))) & 1ULL);
        } else s = margin > 0 ? 0u : 1u;
        sbits |= s << q;
    }
    if (lane < cnt) {
        uint32_t s = (sbits >> lane) & 1u;
        side[my_pos] = (uint8_t)s;
        flag32[my_pos] = (s == 0);
    }
}

// Point-major variant (NND_FOREST_PM, default off): the position-major passes above read every row of X once PER TREE and level
// (rows are gathered in each tree's own permuted order and X does not fit the L2: 4*ld*N*n_trees bytes of DRAM traffic per
// level -- 92 GB per 1M x 128 build with 12 trees).  Here one warp keeps a point's row in registers and walks the point's
// node in EVERY tree: X is read once per level (12x less DRAM traffic), the per-(tree, point) stream becomes the node's
// hyperplane, which comes from the L2 while the level's table (M x ld floats) is small -- the host switches back to the
// position-major kernel when it is not.  Same margin arithmetic (4 fma per lane, xor butterfly, + offset) and the same tie
// hash as k_forest_side128 / k_forest_side, so side[] / flag32[] are bit-identical; invp (maintained by k_forest_scatter) supplies
// position = invp[tree*N + point], -1 for a point that already sits in a final leaf of that tree.  Needs n_trees <= 32 and
// ld <= 1024 (rows of up to eight 128-float chunks in registers: 2M x 256 and 1M x 784 gather 370 and 560 GB per build otherwise).
#define PM_RUN 8
// NCH = number of 128-float chunks a row spans (ld <= 128 * NCH); lane l holds floats [128 ch + 4 l, 128 ch + 4 l + 4) of chunk
// ch -- the very elements warp_lane_dot / k_forest_side give it, accumulated in the same order (chunk by chunk, 4 fma each).
template <int NCH>
__global__ void __launch_bounds__(256) k_forest_side_pm(const float *X, int ld, int64_t N, int n_trees, int level,
                                                        const int32_t *invp, const int32_t *seg, const int32_t *nstart,
                                                        const int32_t *tree_first, const uint64_t *tree_seed, const float *hyper,
                                                        const float *noff, uint8_t *side, uint32_t *flag32) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t p0 = w * PM_RUN;
    if (p0 >= N) return;
    const int cnt = (int)min((int64_t)PM_RUN, N - p0);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 xcur[NCH], xnext[NCH];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
        const int c = ch * 128 + lane * 4;
        xcur[ch] = c < ld ? *(const float4 *)(X + p0 * ld + c) : zero4;
        xnext[ch] = xcur[ch];
    }
    for (int q = 0; q < cnt; ++q) {
        const int64_t p = p0 + q;
        if (q + 1 < cnt) {
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) { const int c = ch * 128 + lane * 4; xnext[ch] = c < ld ? *(const float4 *)(X + (p + 1) * ld + c) : zero4; }
        }
        // lane t < n_trees: where the point sits in tree t and which active node that is
        int32_t my_pos = -1, my_o = -1;
        if (lane < n_trees) { my_pos = invp[(int64_t)lane * N + p]; if (my_pos >= 0) my_o = seg[my_pos]; }
        unsigned live = __ballot_sync(0xffffffffu, my_pos >= 0);
        float4 b[NCH];
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) { b[ch] = xcur[ch]; xcur[ch] = xnext[ch]; }
        uint32_t sbits = 0;
        while (live) {
            const int t = __ffs(live) - 1;
            live &= live - 1;
            const int o = __shfl_sync(0xffffffffu, my_o, t);
            const float *hrow = hyper + (size_t)o * ld;
            float acc = 0.0f;
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                const int c = ch * 128 + lane * 4;
                if (c < ld) {
                    const float4 h4 = *(const float4 *)(hrow + c);
                    acc = __fmaf_rn(h4.x, b[ch].x, acc); acc = __fmaf_rn(h4.y, b[ch].y, acc);
                    acc = __fmaf_rn(h4.z, b[ch].z, acc); acc = __fmaf_rn(h4.w, b[ch].w, acc);
                }
            }
#pragma unroll
            for (int sft = 16; sft >= 1; sft >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, sft));
            const float margin = __fadd_rn(noff[o], acc);
            uint32_t s;
            if (fabs((double)margin) < 1e-8) {
                const int pos = __shfl_sync(0xffffffffu, my_pos, t);
                const int start = nstart[o];
                const int tree = (int)(start / N);
                const uint64_t key = node_key(tree_seed[tree], level, (int64_t)(o - tree_first[tree]));
                s = (uint32_t)(mix64(key ^ (0x100ULL + (uint64_t)(pos - start))) & 1ULL);
            } else s = margin > 0 ? 0u : 1u;
            sbits |= s << t;
        }
        if (my_pos >= 0) {
            const uint32_t s = (sbits >> lane) & 1u;
            side[my_pos] = (uint8_t)s;
            flag32[my_pos] = (s == 0);
        }
    }
}

// degenerate nodes (all points on one side): random split (rp_trees.py:395-403)
__global__ void k_forest_fix_side(int64_t n_act, const int32_t *act, int64_t N, int level, const int32_t *seg,
                                  const int32_t *nstart, const uint8_t *ndegen, const int32_t *tree_first,
                                  const uint64_t *tree_seed, uint8_t *side, uint32_t *flag32) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_act) return;
    const int64_t pos = act[i];
    const int o = seg[pos];
    if (o < 0 || !ndegen[o]) return;
    const int start = nstart[o];
    const int tree = (int)(start / N);
    const uint64_t key = node_key(tree_seed[tree], level, (int64_t)(o - tree_first[tree]));
    uint8_t s = (uint8_t)(mix64(key ^ 0x8000000000000000ULL ^ (uint64_t)(pos - start)) & 1ULL);
    side[pos] = s;
    flag32[pos] = (s == 0);
}

__global__ void k_forest_node(int M, const int32_t *nstart, const int32_t *nsize, const uint32_t *pl, int leaf_size,
                              int depth_left_after, int32_t *nleft, uint8_t *ndegen, uint32_t *nchild, int32_t *scalars,
                              int second_pass) {
    int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= M) return;
    const int start = nstart[o], n = nsize[o];
    int nl = (int)(pl[start + n] - pl[start]);
    if (!second_pass) {
        bool dg = (nl == 0 || nl == n);
        ndegen[o] = dg;
        if (dg) atomicAdd(&scalars[0], 1);
    }
    nleft[o] = nl;
    uint32_t c = 0;
    if (depth_left_after > 0) { c += (nl > leaf_size); c += (n - nl > leaf_size); }
    nchild[o] = c;
}

// next level's node table + segment-start flags
__global__ void k_forest_children(int M, const int32_t *nstart, const int32_t *nsize, const int32_t *nleft,
                                  const uint32_t *nchild_scan, int leaf_size, int depth_left_after, int64_t N,
                                  int32_t *nstart2, int32_t *nsize2, uint8_t *segstart, int32_t *tree_first2) {
    int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= M) return;
    const int start = nstart[o], n = nsize[o], nl = nleft[o];
    if (nl > 0 && nl < n) segstart[start + nl] = 1;
    uint32_t w = nchild_scan[o];
    if (depth_left_after > 0) {
        if (nl > leaf_size) { nstart2[w] = start; nsize2[w] = nl; atomicMin(&tree_first2[start / N], (int32_t)w); ++w; }
        if (n - nl > leaf_size) { nstart2[w] = start + nl; nsize2[w] = n - nl; atomicMin(&tree_first2[(start + nl) / N], (int32_t)w); }
    }
}

// stable partition of every active segment + relabel positions with the child's ordinal; points whose
// child is a final leaf are written straight to leaf_points, the others go on the next active list
__global__ void k_forest_scatter(int64_t n_act, const int32_t *act, const int32_t *perm, const int32_t *seg,
                                 const uint8_t *side, const uint32_t *pl, const int32_t *nstart, const int32_t *nsize,
                                 const int32_t *nleft, const uint32_t *nchild_scan, int leaf_size, int depth_left_after,
                                 int32_t *perm2, int32_t *seg2, int32_t *act2, int32_t *n_act2, int32_t *leaf_points,
                                 uint32_t *flag32, int32_t *invp, int64_t N) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool alive = false; int32_t np32 = 0;
    if (i < n_act) {
        const int64_t pos = act[i];
        const int o = seg[pos];
        const int start = nstart[o], n = nsize[o], nl = nleft[o];
        const int lrank = (int)(pl[pos] - pl[start]);
        const bool left = side[pos] == 0;
        int64_t np = left ? (int64_t)start + lrank : (int64_t)start + nl + ((pos - start) - lrank);
        int32_t child = -1;
        if (depth_left_after > 0) {
            uint32_t w = nchild_scan[o];
            bool lact = nl > leaf_size, ract = (n - nl) > leaf_size;
            if (left) child = lact ? (int32_t)w : -1;
            else child = ract ? (int32_t)(w + (lact ? 1 : 0)) : -1;
        }
        const int32_t pt = perm[pos];
        if (child >= 0) { perm2[np] = pt; seg2[np] = child; alive = true; np32 = (int32_t)np; }
        else leaf_points[np] = pt;
        if (invp) invp[(int64_t)(start / N) * N + pt] = child >= 0 ? (int32_t)np : -1;   // NND_FOREST_PM: point -> next position
        flag32[pos] = 0;  // scan input must be zero outside the (new) active set
    }
    unsigned mk = __ballot_sync(0xffffffffu, alive);
    if (mk) {
        int base = 0;
        if (lane == 0) base = atomicAdd(n_act2, __popc(mk));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (alive) act2[base + __popc(mk & ((1u << lane) - 1))] = np32;
    }
}

__global__ void k_forest_flags(int64_t TN, const uint8_t *segstart, uint32_t *flag32) {
    int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos < TN) flag32[pos] = segstart[pos];
}
__global__ void k_forest_offsets(int64_t TN, const uint8_t *segstart, const uint32_t *pl, int64_t *leaf_offsets) {
    int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos < TN && segstart[pos]) leaf_offsets[pl[pos]] = pos;
    if (pos == TN) leaf_offsets[pl[TN]] = TN;
}
__global__ void k_forest_maxleaf(int64_t n_leaves, const int64_t *leaf_offsets, int32_t *scalars) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_leaves) atomicMax(&scalars[1], (int32_t)(leaf_offsets[i + 1] - leaf_offsets[i]));
}
__global__ void k_fill_i32(int32_t *p, int32_t v, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

static inline uint64_t tree_seed64(const int64_t *s) {
    return mix64(mix64(mix64((uint64_t)s[0]) ^ (uint64_t)s[1]) ^ (uint64_t)s[2]);
}

extern "C" int nnd_forest_build(nnd_ctx *ctx, const float *X, int64_t N, int32_t d, int32_t ld, int32_t angular,
                                const int64_t *tree_rng_states_host, int32_t n_trees, int32_t leaf_size,
                                int32_t max_depth, int32_t *leaf_points, int64_t *leaf_offsets, int64_t *n_leaves_host,
                                int32_t *max_leaf_host, void *ws, size_t ws_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    NND_CHECK(ctx && X && tree_rng_states_host && leaf_points && leaf_offsets && n_leaves_host && max_leaf_host,
              NND_ERR_INVALID, "nnd_forest_build: null argument");
    NND_CHECK(N > 0 && d > 0 && ld >= d && ld % 4 == 0 && n_trees > 0 && leaf_size >= 8, NND_ERR_INVALID,
              "nnd_forest_build: bad shape (N=%lld d=%d ld=%d trees=%d leaf_size=%d)", (long long)N, d, ld, n_trees, leaf_size);
    const int64_t TN = (int64_t)n_trees * N;
    NND_CHECK(TN < (1LL << 31) - 64, NND_ERR_UNSUPPORTED, "n_trees*N must be < 2^31");
    ForestWs W; size_t used = 0;
    NND_CHECK(carve_forest_ws(ws, ws_bytes, N, ld, n_trees, &W, &used) == NND_OK, NND_ERR_WORKSPACE, "forest: workspace too small");
    used = (used + 255) & ~size_t(255);
    const int64_t hyper_nodes = TN / (leaf_size + 1) + n_trees + 16;
    NND_CHECK(used + (size_t)hyper_nodes * ld * 4 <= ws_bytes, NND_ERR_WORKSPACE, "forest: workspace too small for hyperplanes");
    W.hyper = (float *)((char *)ws + used);

    std::vector<uint64_t> seeds(n_trees);
    for (int t = 0; t < n_trees; ++t) seeds[t] = tree_seed64(tree_rng_states_host + 3 * t);
    NND_CUDA(cudaMemcpyAsync(W.tree_seed, seeds.data(), sizeof(uint64_t) * n_trees, cudaMemcpyHostToDevice, st));
    const bool root_active = (N > leaf_size && max_depth > 0);
    const unsigned gpos = (unsigned)cdiv(TN + 1, 256);
    // NND_FOREST_PM: the inverse permutation is maintained (init + scatter) only when the point-major pass can be used at all
    const bool pm_on = ctx->opt_forest_pm && ld <= 1024 && n_trees <= 32;
    int32_t *invp = pm_on ? W.invp : nullptr;
    NND_STAGE(ctx, NND_ST_FOREST_PART, st);
    k_forest_init<<<gpos, 256, 0, st>>>(W.perm[0], W.seg[0], W.segstart, W.act[0], leaf_points, W.flag32, N, n_trees, root_active ? 1 : 0, invp); NND_LAUNCH(ctx);
    NND_CUDA(cudaMemsetAsync(W.scalars, 0, sizeof(int32_t) * 16, st));
    int cur = 0;
    int M = 0;
    if (root_active) {
        std::vector<int32_t> hs(n_trees), hz(n_trees), hf(n_trees);
        for (int t = 0; t < n_trees; ++t) { hs[t] = (int32_t)((int64_t)t * N); hz[t] = (int32_t)N; hf[t] = t; }
        NND_CUDA(cudaMemcpyAsync(W.nstart[0], hs.data(), sizeof(int32_t) * n_trees, cudaMemcpyHostToDevice, st));
        NND_CUDA(cudaMemcpyAsync(W.nsize[0], hz.data(), sizeof(int32_t) * n_trees, cudaMemcpyHostToDevice, st));
        NND_CUDA(cudaMemcpyAsync(W.tree_first, hf.data(), sizeof(int32_t) * n_trees, cudaMemcpyHostToDevice, st));
        NND_CUDA(cudaStreamSynchronize(st));  // host vectors go out of scope
        M = n_trees;
    }
    int ncur = 0;
    int64_t n_act = root_active ? TN : 0;
    for (int level = 0; M > 0; ++level) {
        NND_CHECK(M <= hyper_nodes, NND_ERR_OVERFLOW, "forest: %d active nodes exceed the hyperplane pool", M);
        const int depth_left_after = max_depth - (level + 1);
        const int hw = 4;
        NND_STAGE(ctx, NND_ST_FOREST_HYPER, st);
        k_forest_hyper<<<(unsigned)cdiv(M, hw), hw * 32, (size_t)hw * ld * 4, st>>>(X, ld, N, angular, level, M, W.nstart[ncur],
                                                                                 W.nsize[ncur], W.perm[cur], W.tree_first,
                                                                                 W.tree_seed, W.hyper, W.noff);
        NND_LAUNCH(ctx);
        NND_STAGE(ctx, NND_ST_FOREST_SIDE, st);
        if (pm_on && (size_t)M * ld * 4 <= (size_t)ctx->opt_forest_pm_mb << 20) {
            const unsigned g = (unsigned)cdiv(cdiv(N, PM_RUN) * 32, 256);
#define PMK(NCH_) k_forest_side_pm<NCH_><<<g, 256, 0, st>>>(X, ld, N, n_trees, level, invp, W.seg[cur], W.nstart[ncur], W.tree_first, \
                                                          W.tree_seed, W.hyper, W.noff, W.side, W.flag32)
            if (ld <= 128) PMK(1); else if (ld <= 256) PMK(2); else if (ld <= 512) PMK(4); else PMK(8);
#undef PMK
        } else if (ld <= 128)
            k_forest_side128<<<(unsigned)cdiv(cdiv(n_act, SIDE_RUN) * 32, 256), 256, 0, st>>>(X, ld, n_act, W.act[cur], N, level,
                                                                                         W.perm[cur], W.seg[cur], W.nstart[ncur],
                                                                                         W.tree_first, W.tree_seed, W.hyper, W.noff,
                                                                                         W.side, W.flag32);
        else
            k_forest_side<<<(unsigned)cdiv(cdiv(n_act, SIDE_PW) * 32, 256), 256, 0, st>>>(X, ld, n_act, W.act[cur], N, level, W.perm[cur],
                                                                                       W.seg[cur], W.nstart[ncur], W.tree_first,
                                                                                       W.tree_seed, W.hyper, W.noff, W.side, W.flag32);
        NND_LAUNCH(ctx);
        NND_STAGE(ctx, NND_ST_FOREST_PART, st);
        NND_TRY(nnd_exclusive_scan_u32(ctx, W.flag32, W.pl, TN, W.scan_tmp, st));
        NND_CUDA(cudaMemsetAsync(W.scalars, 0, sizeof(int32_t), st));
        NND_CUDA(cudaMemsetAsync(W.scalars + 2, 0, sizeof(int32_t), st));
        k_forest_node<<<(unsigned)cdiv(M, 256), 256, 0, st>>>(M, W.nstart[ncur], W.nsize[ncur], W.pl, leaf_size, depth_left_after,
                                                              W.nleft, W.ndegen, W.nchild, W.scalars, 0);
        NND_LAUNCH(ctx);
        int32_t hsc[2];
        NND_CUDA(cudaMemcpyAsync(hsc, W.scalars, sizeof(int32_t) * 2, cudaMemcpyDeviceToHost, st));
        NND_CUDA(cudaStreamSynchronize(st));
        if (hsc[0] > 0) {
            k_forest_fix_side<<<(unsigned)cdiv(n_act, 256), 256, 0, st>>>(n_act, W.act[cur], N, level, W.seg[cur], W.nstart[ncur],
                                                                          W.ndegen, W.tree_first, W.tree_seed, W.side, W.flag32);
            NND_LAUNCH(ctx);
            NND_TRY(nnd_exclusive_scan_u32(ctx, W.flag32, W.pl, TN, W.scan_tmp, st));
            k_forest_node<<<(unsigned)cdiv(M, 256), 256, 0, st>>>(M, W.nstart[ncur], W.nsize[ncur], W.pl, leaf_size,
                                                                  depth_left_after, W.nleft, W.ndegen, W.nchild, W.scalars, 1);
            NND_LAUNCH(ctx);
        }
        NND_TRY(nnd_exclusive_scan_u32(ctx, W.nchild, W.nchild_scan, M, W.scan_tmp, st));
        k_fill_i32<<<1, 256, 0, st>>>(W.tree_first, 0x7fffffff, n_trees); NND_LAUNCH(ctx);
        k_forest_children<<<(unsigned)cdiv(M, 256), 256, 0, st>>>(M, W.nstart[ncur], W.nsize[ncur], W.nleft, W.nchild_scan,
                                                                  leaf_size, depth_left_after, N, W.nstart[ncur ^ 1],
                                                                  W.nsize[ncur ^ 1], W.segstart, W.tree_first);
        NND_LAUNCH(ctx);
        k_forest_scatter<<<(unsigned)cdiv(n_act, 256), 256, 0, st>>>(n_act, W.act[cur], W.perm[cur], W.seg[cur], W.side, W.pl,
                                                                     W.nstart[ncur], W.nsize[ncur], W.nleft, W.nchild_scan, leaf_size,
                                                                     depth_left_after, W.perm[cur ^ 1], W.seg[cur ^ 1], W.act[cur ^ 1],
                                                                     W.scalars + 2, leaf_points, W.flag32, invp, N);
        NND_LAUNCH(ctx);
        uint32_t m_next = 0; int32_t a_next = 0;
        NND_CUDA(cudaMemcpyAsync(&m_next, W.nchild_scan + M, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        NND_CUDA(cudaMemcpyAsync(&a_next, W.scalars + 2, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        NND_CUDA(cudaStreamSynchronize(st));
        n_act = a_next;
        cur ^= 1; ncur ^= 1;
        M = (int)m_next;
    }
    // leaves = final segments in position order
    k_forest_flags<<<(unsigned)cdiv(TN, 256), 256, 0, st>>>(TN, W.segstart, W.flag32); NND_LAUNCH(ctx);
    NND_TRY(nnd_exclusive_scan_u32(ctx, W.flag32, W.pl, TN, W.scan_tmp, st));
    k_forest_offsets<<<gpos, 256, 0, st>>>(TN, W.segstart, W.pl, leaf_offsets); NND_LAUNCH(ctx);
    uint32_t nl = 0;
    NND_CUDA(cudaMemcpyAsync(&nl, W.pl + TN, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    NND_CUDA(cudaStreamSynchronize(st));
    NND_CUDA(cudaMemsetAsync(W.scalars + 1, 0, sizeof(int32_t), st));
    k_forest_maxleaf<<<(unsigned)cdiv((int64_t)nl, 256), 256, 0, st>>>((int64_t)nl, leaf_offsets, W.scalars); NND_LAUNCH(ctx);
    int32_t ml = 0;
    NND_CUDA(cudaMemcpyAsync(&ml, W.scalars + 1, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    NND_CUDA(cudaStreamSynchronize(st));
    *n_leaves_host = (int64_t)nl;
    *max_leaf_host = ml > leaf_size ? ml : leaf_size;
    NND_STAGE(ctx, NND_ST_OTHER, st);
    NND_CUDA(cudaGetLastError());
    return NND_OK;
}
