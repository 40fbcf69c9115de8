// k_cand.cu -- new_build_candidates (utils.py:221-320) for T emulated numba threads.
//
// Reference: thread n owns vertex range R_n=[n*B,(n+1)*B), B=N//T+1, uses stream rng_state+n,
// scans ALL edges (i,j) in heap-layout order and draws ONE priority per edge that touches R_n
// (utils.py:266-306); the edge is pushed into cand[flag][i] (if i in R_n) and cand[flag][idx]
// (if idx in R_n) with checked_heap_push.  Then flags of entries present in new_cand[i] are
// cleared (utils.py:311-318).
//
// GPU formulation (bit-exact):
//   k_cand_count : per row, #valid edges and #edges pointing into each R_n; in-degree histogram
//   scan         : prefix sums -> the draw index of every (edge, stream) pair is a closed form
//   k_cand_fill  : priority = table[n][draw index] (tables = the T streams, generated once by
//                  GF(2) jump-ahead); forward priorities stored per edge, reverse edges routed
//                  to their target vertex (counting sort by target)
//   k_cand_select: one warp per vertex sorts its reverse edges by scan position, merges them with
//                  its forward edges and replays checked_heap_push in scan order; clears flags
#include "nnd_common.cuh"
#include <vector>

struct CandRec;
struct CandWs {
    uint32_t *cnt;      // [(T+1)*N] per-stream per-row counts; last block = valid counts
    uint32_t *pre;      // [(T+1)*N + 1] exclusive scan of cnt
    uint32_t *indeg;    // [N+1]
    uint32_t *revoff;   // [N+2] padded-segment offsets
    uint32_t *cursor;   // [N]
    unsigned long long *rev;  // [2*N*k + pad]  (key<<32 | prio bits), key = (i*k+j)*2 + isnew
    float *fprio;       // [N*k]
    float *table;       // [T * table_len]
    uint32_t *scan_tmp;
    int64_t table_len;
    uint8_t *act;       // [N+1] NND_CAND_SKIP: 1 = the vertex touches a new-flagged edge (= will have new candidates)
    unsigned long long *scal;   // [8192] small tables of the partitioned path (layout: SC_* below)
    struct CandRec *recs;       // [N*k] reverse-edge records of the single-process partitioned build
};

static int64_t cand_table_len(int64_t N, int k) { return N * (int64_t)k; }

static int carve_cand_ws(void *ws, size_t ws_bytes, int64_t N, int k, int C, int T, CandWs *W, size_t *need) {
    WsAlloc A(ws, ws_bytes);
    (void)C;
    int64_t Nb = N > 16 ? N : 16;
    W->cnt = A.get<uint32_t>((size_t)(T + 1) * Nb);
    W->pre = A.get<uint32_t>((size_t)(T + 1) * Nb + 2);
    W->indeg = A.get<uint32_t>(Nb + 1);
    W->revoff = A.get<uint32_t>(Nb + 2);
    W->cursor = A.get<uint32_t>(Nb + 1);
    W->rev = A.get<unsigned long long>((size_t)2 * Nb * k + NND_SEG_SMEM);
    W->fprio = A.get<float>((size_t)Nb * k);
    W->table_len = cand_table_len(Nb, k);
    W->table = A.get<float>((size_t)T * W->table_len);
    W->scan_tmp = A.get<uint32_t>(nnd_scan_tmp_elems((int64_t)(T + 1) * Nb + 2));
    W->act = A.get<uint8_t>(Nb + 1);
    W->scal = A.get<unsigned long long>(8192);
    W->recs = (CandRec *)A.get<char>((size_t)Nb * k * 12);
    if (need) *need = A.off;
    return A.ok ? NND_OK : NND_ERR_WORKSPACE;
}
extern "C" int nnd_cand_workspace_bytes(int64_t N, int32_t k, int32_t C, int32_t T, size_t *bytes) {
    NND_CHECK(bytes && N > 0 && k > 0 && C > 0 && T > 0, NND_ERR_INVALID, "nnd_cand_workspace_bytes: bad args");
    *bytes = nnd_cand_ws_bytes(N, k, C, T);
    return NND_OK;
}
size_t nnd_cand_ws_bytes(int64_t N, int k, int C, int T) {
    CandWs W; size_t need = 0;
    carve_cand_ws(nullptr, 0, N, k, C, T, &W, &need);
    return need + 1024;
}

// act (NND_CAND_SKIP, may be null): a vertex gets at least one NEW candidate iff it is an endpoint of a new-flagged edge
// (the first push into its empty new-heap is always accepted, utils.py:289-306), and the local join skips every vertex
// without new candidates (utils.py:575-577 loops over new candidates only) -- so the candidate lists of all other vertices
// are never read and need not be sampled.
__global__ void k_cand_count(const int32_t *ind, const uint8_t *flags, int k, int64_t N, int T, int64_t B, uint32_t *cnt,
                             uint32_t *indeg, uint8_t *act) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    uint32_t valid = 0;
    for (int n = 0; n < T; ++n) {
        int64_t bs = (int64_t)n * B, be = bs + B;
        uint32_t c = 0;
        for (int j = 0; j < k; ++j) { int32_t x = ind[i * k + j]; c += (x >= bs && x < be && x >= 0); }
        cnt[(size_t)n * N + i] = c;
    }
    bool anynew = false;
    for (int j = 0; j < k; ++j) {
        int32_t x = ind[i * k + j];
        if (x >= 0) {
            ++valid; atomicAdd(&indeg[x], 1u);
            if (act && flags[i * k + j]) { anynew = true; act[x] = 1; }
        }
    }
    if (anynew) act[i] = 1;
    cnt[(size_t)T * N + i] = valid;
}

// draw index of edge (i, j) in stream n: number of earlier edges (scan order) touching R_n
__device__ __forceinline__ uint32_t draw_index(const uint32_t *pre, int64_t N, int T, int64_t B, int n, int64_t i,
                                               uint32_t within_valid, uint32_t within_n) {
    const uint32_t *PR = pre + (size_t)n * N;  // prefix of "#edges into R_n" per row (relative to PR[0] base)
    const uint32_t *PV = pre + (size_t)T * N;  // prefix of valid counts
    const uint32_t baseR = PR[0], baseV = PV[0];
    int64_t s = (int64_t)n * B, e = s + B < N ? s + B : N;
    // PR[x]-baseR = sum over rows < x ; index N is not addressable inside the block, handle via next block's base
    auto prR = [&](int64_t x) -> uint32_t { return (x < N ? PR[x] : pre[(size_t)(n + 1) * N]) - baseR; };
    auto prV = [&](int64_t x) -> uint32_t { return (x < N ? PV[x] : pre[(size_t)(T + 1) * N]) - baseV; };
    if (i < s) return prR(i) + within_n;
    if (i < e) return prR(s) + (prV(i) - prV(s)) + within_valid;
    return prR(s) + (prV(e) - prV(s)) + (prR(i) - prR(e)) + within_n;
}

__global__ void k_cand_fill(const int32_t *ind, const uint8_t *flags, int k, int64_t N, int T, int64_t B,
                            const uint32_t *pre, const float *table, int64_t table_len, const uint32_t *revoff,
                            uint32_t *cursor, unsigned long long *rev, float *fprio, int64_t lo, int64_t hi,
                            const uint8_t *act) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const bool mine = i >= lo && i < hi;
    const int ni = (int)(i / B);
    const bool ai = !act || act[i];
    uint32_t wv = 0;  // valid edges before j in this row
    for (int j = 0; j < k; ++j) {
        int32_t x = ind[i * k + j];
        if (x < 0) { fprio[i * k + j] = 0.f; continue; }
        if (!mine && (x < lo || x >= hi)) { ++wv; continue; }   // neither endpoint sampled by this process
        if (act) {
            // NND_CAND_SKIP: the forward priority is only read by select(i), the reverse entry only by select(x); both are
            // skipped for vertices without new candidates (draw indices are closed forms, nothing else depends on them)
            const bool ax = act[x] && x >= lo && x < hi;
            if (!(ai && mine) && !ax) { ++wv; continue; }
        }
        const int nx = (int)(x / B);
        uint32_t df = draw_index(pre, N, T, B, ni, i, wv, 0);
        float pf = table[(size_t)ni * table_len + df];
        fprio[i * k + j] = pf;
        float pr = pf;
        if (nx != ni) {
            uint32_t wn = 0;  // earlier edges of this row pointing into R_nx
            int64_t bs = (int64_t)nx * B, be = bs + B;
            for (int jj = 0; jj < j; ++jj) { int32_t y = ind[i * k + jj]; wn += (y >= bs && y < be); }
            uint32_t dr = draw_index(pre, N, T, B, nx, i, 0, wn);
            pr = table[(size_t)nx * table_len + dr];
        }
        if (x < lo || x >= hi) { ++wv; continue; }
        if (act && !act[x]) { ++wv; continue; }
        uint32_t slot = atomicAdd(&cursor[x], 1u);
        unsigned long long key = ((unsigned long long)(i * k + j) << 1) | (unsigned long long)(flags[i * k + j] != 0);
        rev[(size_t)revoff[x] + slot] = (key << 32) | (unsigned long long)__float_as_uint(pr);
        ++wv;
    }
}

#define SEL_SMALL 32
#define SEL_WARPS 8
__global__ void __launch_bounds__(SEL_WARPS * 32) k_cand_select(const int32_t *ind, uint8_t *flags, int k, int64_t N, int C,
                                                                const uint32_t *indeg, const uint32_t *revoff,
                                                                unsigned long long *rev, const float *fprio,
                                                                int32_t *newc, int32_t *oldc, int skip_small, int64_t lo, int64_t hi,
                                                                const uint8_t *act) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t v = (int64_t)blockIdx.x * SEL_WARPS + warp;
    if (v >= N || v < lo || v >= hi) return;
    unsigned long long *buf = (unsigned long long *)smem_raw + (size_t)warp * NND_SEG_SMEM;
    const size_t hstride = ((size_t)C * 16 + 15) & ~size_t(15);
    unsigned char *hp = smem_raw + (size_t)SEL_WARPS * NND_SEG_SMEM * 8 + (size_t)warp * hstride;
    float *npri = (float *)hp, *opri = npri + C;
    int32_t *nind = (int32_t *)(opri + C), *oind = nind + C;
    for (int c = lane; c < C; c += 32) { npri[c] = CUDART_INF_F; opri[c] = CUDART_INF_F; nind[c] = -1; oind[c] = -1; }
    const uint32_t nrev = indeg[v];
    if (skip_small && nrev <= SEL_SMALL) return;
    if (act && !act[v]) {     // no new candidates: the join never looks at this vertex
        for (int c = lane; c < C; c += 32) { newc[v * C + c] = -1; oldc[v * C + c] = -1; }
        return;
    }
    unsigned long long *seg = rev + revoff[v];
    unsigned long long *sorted;
    if (nrev <= NND_SEG_SMEM) {
        int p2 = (int)next_pow2_u32(nrev);
        for (int j = lane; j < p2; j += 32) buf[j] = (j < (int)nrev) ? seg[j] : ~0ULL;
        __syncwarp();
        if (nrev > 1) warp_bitonic_sort_u64(buf, p2, lane);
        sorted = buf;
    } else {
        uint32_t cap = padded_seg(nrev);
        for (uint32_t j = nrev + lane; j < cap; j += 32) seg[j] = ~0ULL;
        __syncwarp();
        warp_bitonic_sort_u64(seg, (int)cap, lane);
        sorted = seg;
    }
    __syncwarp();
    // two-pointer merge of forward edges (v*k+j, ascending j) and sorted reverse edges
    int jf = 0; uint32_t jr = 0;
    const unsigned long long fbase = (unsigned long long)v * k;
    for (;;) {
        while (jf < k && ind[v * k + jf] < 0) ++jf;
        bool hasf = jf < k, hasr = jr < nrev;
        if (!hasf && !hasr) break;
        unsigned long long rkey = hasr ? (sorted[jr] >> 33) : ~0ULL;  // scan position of the reverse edge
        bool take_f = hasf && (!hasr || fbase + jf <= rkey);
        float pr; int32_t cand; bool isnew;
        if (take_f) {
            pr = fprio[v * k + jf]; cand = ind[v * k + jf]; isnew = flags[v * k + jf] != 0; ++jf;
        } else {
            unsigned long long r = sorted[jr]; ++jr;
            pr = __uint_as_float((uint32_t)(r & 0xFFFFFFFFULL));
            isnew = (r >> 32) & 1ULL;
            cand = (int32_t)((r >> 33) / (unsigned long long)k);
        }
        if (isnew) warp_heap_push(npri, nind, nullptr, C, pr, cand, 0, lane);
        else warp_heap_push(opri, oind, nullptr, C, pr, cand, 0, lane);
    }
    __syncwarp();
    // utils.py:311-318: clear the flag of every neighbour that was sampled into new_cand[v]
    for (int j = lane; j < k; j += 32) {
        int32_t x = ind[v * k + j];
        bool found = false;
        for (int c = 0; c < C; ++c) found |= (nind[c] == x);
        if (found) flags[v * k + j] = 0;
    }
    for (int c = lane; c < C; c += 32) { newc[v * C + c] = nind[c]; oldc[v * C + c] = oind[c]; }
}


// ---- thread-per-vertex variant for vertices with <= SEL_SMALL reverse edges (the vast majority):
// the reverse edges are staged in shared memory (column per thread), the next edge in scan order
// is found by selection, and checked_heap_push is replayed by the thread on column-layout heaps.
#define SEL_T 128
__global__ void __launch_bounds__(SEL_T) k_cand_select_small(const int32_t *ind, uint8_t *flags, int k, int64_t N, int C,
                                                             const uint32_t *indeg, const uint32_t *revoff,
                                                             const unsigned long long *rev, const float *fprio,
                                                             int32_t *newc, int32_t *oldc, int64_t lo, int64_t hi,
                                                             const uint8_t *act) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int64_t v = (int64_t)blockIdx.x * SEL_T + tid;
    float *npri = (float *)smem_raw + tid;                                         // [C][SEL_T]
    float *opri = npri + (size_t)C * SEL_T;
    int32_t *nind = (int32_t *)(opri + (size_t)C * SEL_T), *oind = nind + (size_t)C * SEL_T;
    if (v >= N || v < lo || v >= hi) return;
    const uint32_t nrev = indeg[v];
    if (nrev > SEL_SMALL) return;  // handled by the warp kernel
    if (act && !act[v]) {          // no new candidates: the join never looks at this vertex
        for (int c = 0; c < C; ++c) { newc[v * C + c] = -1; oldc[v * C + c] = -1; }
        return;
    }
    for (int c = 0; c < C; ++c) { npri[c * SEL_T] = CUDART_INF_F; opri[c * SEL_T] = CUDART_INF_F; nind[c * SEL_T] = -1; oind[c * SEL_T] = -1; }
    const unsigned long long *seg = rev + revoff[v];
    int jf = 0; uint32_t left = nrev, used = 0;   // bit j of `used`: reverse edge j already consumed
    const unsigned long long fbase = (unsigned long long)v * k;
    for (;;) {
        while (jf < k && ind[v * k + jf] < 0) ++jf;
        const bool hasf = jf < k;
        // smallest remaining reverse entry (keys are unique)
        unsigned long long best = ~0ULL; int bj = -1;
        if (left)
            {   // select on the 32-bit (scan position, isnew) word only; the priority word is fetched for the winner
                const uint32_t *kw = (const uint32_t *)seg;
                uint32_t bk = 0xFFFFFFFFu;
                for (uint32_t j = 0; j < nrev; ++j) {
                    if ((used >> j) & 1u) continue;
                    uint32_t x = kw[2 * j + 1];
                    if (x < bk) { bk = x; bj = (int)j; }
                }
                if (bj >= 0) best = seg[bj];
            }
        const bool hasr = bj >= 0;
        if (!hasf && !hasr) break;
        const bool take_f = hasf && (!hasr || fbase + jf <= (best >> 33));
        float p; int32_t cand; bool isnew;
        if (take_f) { p = fprio[v * k + jf]; cand = ind[v * k + jf]; isnew = flags[v * k + jf] != 0; ++jf; }
        else {
            used |= 1u << bj; --left;
            p = __uint_as_float((uint32_t)(best & 0xFFFFFFFFULL)); isnew = (best >> 32) & 1ULL;
            cand = (int32_t)((best >> 33) / (unsigned long long)k);
        }
        float *pri = isnew ? npri : opri; int32_t *idx = isnew ? nind : oind;
        if (p >= pri[0]) continue;
        bool dup = false;
        for (int c = 0; c < C; ++c) dup |= (idx[c * SEL_T] == cand);
        if (dup) continue;
        int i = 0, i_swap;
        for (;;) {
            int ic1 = 2 * i + 1, ic2 = ic1 + 1;
            if (ic1 >= C) break;
            else if (ic2 >= C) { if (pri[ic1 * SEL_T] > p) i_swap = ic1; else break; }
            else if (pri[ic1 * SEL_T] >= pri[ic2 * SEL_T]) { if (p < pri[ic1 * SEL_T]) i_swap = ic1; else break; }
            else { if (p < pri[ic2 * SEL_T]) i_swap = ic2; else break; }
            pri[i * SEL_T] = pri[i_swap * SEL_T]; idx[i * SEL_T] = idx[i_swap * SEL_T];
            i = i_swap;
        }
        pri[i * SEL_T] = p; idx[i * SEL_T] = cand;
    }
    for (int j = 0; j < k; ++j) {
        int32_t x = ind[v * k + j];
        bool found = false;
        for (int c = 0; c < C; ++c) found |= (nind[c * SEL_T] == x);
        if (found) flags[v * k + j] = 0;
    }
    for (int c = 0; c < C; ++c) { newc[v * C + c] = nind[c * SEL_T]; oldc[v * C + c] = oind[c * SEL_T]; }
}

__global__ void k_rng_table(const uint64_t *jump, long long s0, long long s1, long long s2, int64_t start, int64_t count,
                            int as_float, void *out, int chunk) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t b = t * chunk;
    if (b >= count) return;
    long long s[3] = {s0, s1, s2};
    tau_jump(s, (unsigned long long)(start + b), jump);
    int64_t e = b + chunk < count ? b + chunk : count;
    for (int64_t i = b; i < e; ++i) {
        int32_t r = tau_step(s);
        if (as_float) ((float *)out)[i] = tau_to_float(r);
        else ((int32_t *)out)[i] = r;
    }
}

int nnd_rng_table_launch(nnd_ctx *ctx, const int64_t *state, int64_t start, int64_t count, int as_float, void *out,
                         cudaStream_t st) {
    if (count <= 0) return NND_OK;
    const int chunk = 256;
    int64_t nthreads = cdiv(count, chunk);
    k_rng_table<<<(unsigned)cdiv(nthreads, 128), 128, 0, st>>>(ctx->d_jump, state[0], state[1], state[2], start, count,
                                                              as_float, out, chunk);
    NND_LAUNCH(ctx);
    NND_CUDA(cudaGetLastError());
    return NND_OK;
}

extern "C" int nnd_rng_stream(nnd_ctx *ctx, const int64_t *state_host, int64_t start, int64_t count, int32_t as_float,
                              void *out, void *stream) {
    NND_CHECK(ctx && state_host && out && start >= 0 && count >= 0, NND_ERR_INVALID, "nnd_rng_stream: bad args");
    return nnd_rng_table_launch(ctx, state_host, start, count, as_float, out, (cudaStream_t)stream);
}

// heap selection + flag clearing for the vertices of [ctx->row_lo, ctx->row_hi) from the routed reverse edges (W.rev) and the
// forward priorities (W.fprio): the tail of nnd_build_candidates, shared with the sharded import path
static int launch_cand_select(nnd_ctx *ctx, const nnd_problem *P, int32_t C, CandWs &W, int32_t *new_cand, int32_t *old_cand,
                              uint8_t *act, cudaStream_t st) {
    const int64_t N = P->N; const int k = P->k;
    const size_t hstride = ((size_t)C * 16 + 15) & ~size_t(15);
    size_t smem = (size_t)SEL_WARPS * NND_SEG_SMEM * 8 + (size_t)SEL_WARPS * hstride;
    NND_CHECK(smem <= 200 * 1024, NND_ERR_UNSUPPORTED, "cand_select: C=%d too large", C);
    size_t smem_small = (size_t)C * SEL_T * 16;
    const int use_small = smem_small <= 200 * 1024;
    if (ctx->opt_carveout) {   // NND_SMEM_CARVEOUT=1: ask for the largest shared-memory carveout (ncu: select_small ran 3 CTAs/SM without)
        cudaFuncSetAttribute(k_cand_select_small, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_cand_select, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    if (use_small) {
        NND_CUDA(cudaFuncSetAttribute(k_cand_select_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_small));
        k_cand_select_small<<<(unsigned)cdiv(N, SEL_T), SEL_T, smem_small, st>>>(P->heap_idx, P->heap_flag, k, N, C, W.indeg, W.revoff, W.rev, W.fprio,
                                                                            new_cand, old_cand, ctx->row_lo, ctx->row_hi, act);
        NND_LAUNCH(ctx);
    }
    NND_CUDA(cudaFuncSetAttribute(k_cand_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_cand_select<<<(unsigned)cdiv(N, SEL_WARPS), SEL_WARPS * 32, smem, st>>>(P->heap_idx, P->heap_flag, k, N, C, W.indeg,
                                                                             W.revoff, W.rev, W.fprio, new_cand, old_cand, use_small, ctx->row_lo, ctx->row_hi, act);
    NND_LAUNCH(ctx);
    return NND_OK;
}

struct CandWs;
int nnd_build_candidates_partitioned(nnd_ctx *ctx, const nnd_problem *P, int32_t C, const int64_t *rng_state_host, int32_t T, int32_t tables_valid,
                                     int32_t *new_cand, int32_t *old_cand, CandWs &W, cudaStream_t st);
extern "C" int nnd_build_candidates(nnd_ctx *ctx, const nnd_problem *P, int32_t C, const int64_t *rng_state_host,
                                    int32_t T, int32_t tables_valid, int32_t *new_cand, int32_t *old_cand, void *ws,
                                    size_t ws_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    NND_CHECK(ctx && P && rng_state_host && new_cand && old_cand, NND_ERR_INVALID, "nnd_build_candidates: bad args");
    NND_CHECK(T >= 1 && T <= 256 && C >= 1 && C <= 256, NND_ERR_INVALID, "nnd_build_candidates: T=%d C=%d out of range", T, C);
    const int64_t N = P->N; const int k = P->k;
    NND_CHECK((double)N * k * 2 < 4.0e9, NND_ERR_UNSUPPORTED, "N*k too large for 32-bit edge keys");
    CandWs W;
    NND_CHECK(carve_cand_ws(ws, ws_bytes, N, k, C, T, &W, nullptr) == NND_OK, NND_ERR_WORKSPACE,
              "nnd_build_candidates: workspace too small");
    const int64_t B = N / T + 1;
    NND_STAGE(ctx, NND_ST_CAND_FILL, st);
    if (ctx->opt_cand_partition && ctx->row_lo <= 0 && ctx->row_hi >= N) {
        NND_TRY(nnd_build_candidates_partitioned(ctx, P, C, rng_state_host, T, tables_valid, new_cand, old_cand, W, st));
        NND_STAGE(ctx, NND_ST_OTHER, st);
        NND_CUDA(cudaGetLastError());
        return NND_OK;
    }
    // NND_CAND_PARTITION=0: the direct scatter (round-1 path, kept for the A/B)
    // the T priority streams: constant across iterations (utils.py:262 copies rng_state, never advances it)
    if (!tables_valid) {
        for (int n = 0; n < T; ++n) {
            int64_t s[3] = {rng_state_host[0] + n, rng_state_host[1] + n, rng_state_host[2] + n};
            NND_TRY(nnd_rng_table_launch(ctx, s, 0, W.table_len, 1, W.table + (size_t)n * W.table_len, st));
        }
    }
    NND_CUDA(cudaMemsetAsync(W.indeg, 0, sizeof(uint32_t) * (N + 1), st));
    NND_CUDA(cudaMemsetAsync(W.cursor, 0, sizeof(uint32_t) * (N + 1), st));
    uint8_t *act = ctx->opt_cand_skip ? W.act : nullptr;
    if (act) NND_CUDA(cudaMemsetAsync(act, 0, (size_t)N + 1, st));
    k_cand_count<<<(unsigned)cdiv(N, 128), 128, 0, st>>>(P->heap_idx, P->heap_flag, k, N, T, B, W.cnt, W.indeg, act); NND_LAUNCH(ctx);
    NND_TRY(nnd_exclusive_scan_u32(ctx, W.cnt, W.pre, (int64_t)(T + 1) * N, W.scan_tmp, st));
    NND_TRY(nnd_exclusive_scan_u32_padded(ctx, W.indeg, W.revoff, N, W.scan_tmp, st));
    k_cand_fill<<<(unsigned)cdiv(N, 128), 128, 0, st>>>(P->heap_idx, P->heap_flag, k, N, T, B, W.pre, W.table, W.table_len,
                                                       W.revoff, W.cursor, W.rev, W.fprio, ctx->row_lo, ctx->row_hi, act);
    NND_LAUNCH(ctx);
    NND_STAGE(ctx, NND_ST_CAND_SELECT, st);
    NND_TRY(launch_cand_select(ctx, P, C, W, new_cand, old_cand, act, st));
    NND_STAGE(ctx, NND_ST_OTHER, st);
    NND_CUDA(cudaGetLastError());
    return NND_OK;
}

// =====================================================================================================================
// Partitioned candidate build (the default on one GPU and the row-sharded build's only path, DESIGN.md sections 6 and 8).
// A process samples candidates for its own rows [row_lo, row_hi) only and SCANS only its own rows' edges (the reference's
// threads each scan all N*k edges, utils.py:261-306).  Reverse edges are not scattered straight to their target (at 10M points
// that is four DRAM-random 4..8-byte accesses per edge: 128 ms per build) but APPENDED to the bucket of the target's range
// (<= 1024 buckets of ~10k consecutive targets, one sequential write stream each); the import pass then reads the records in
// memory order, so that the per-target counters, offsets and segments it touches at any time fit the L2.  Buckets are grouped by
// owner process, so the same records are what crosses NVLink in the sharded build:
//   * a small table of per-stream counts (all-gather by the caller): the draw index of an edge in stream n is the number of
//     earlier edges, in (row, column) order, that touch R_n -- a prefix over ALL rows, i.e. sum over lower ranks' totals + a local
//     prefix; the four prefixes at the stream's own range boundaries s_n / e_n are sums over the processes' partial prefixes;
//   * the reverse-edge proposals (target, scan key, priority) of edges whose target row lives elsewhere (one all-to-all-v).
// Results are identical to the reference's new_build_candidates for any number of processes and buckets (closed-form draw
// indices, ordered replay).
// =====================================================================================================================
struct __align__(4) CandRec { int32_t tgt; uint32_t key; float prio; };   // 12 bytes on the wire
#define CAND_MAXB 1024
// scal layout (u64 words): [0,1024) bucket counts, [1024,2048) bucket offsets, [2048,3072) bucket cursors, [3072,3136) totals per
// owner process, [3136] all records, [4096,4096+5T+1] this process's count table, [6144,..) the global count table (u32)
#define SC_CNT 0
#define SC_OFF 1024
#define SC_CUR 2048
#define SC_TOT 3072
#define SC_ALL 3136
#define SC_LOC 4096
#define SC_GLOB 6144

struct BucketMap { int64_t owner_block; int R; int nbpr; int64_t bsize; };
static BucketMap bucket_map(int64_t owner_block, int R) {
    BucketMap m; m.owner_block = owner_block; m.R = R;
    m.nbpr = CAND_MAXB / R; if (m.nbpr < 1) m.nbpr = 1;
    m.bsize = (owner_block + m.nbpr - 1) / m.nbpr; if (m.bsize < 1) m.bsize = 1;
    return m;
}
__device__ __forceinline__ int bucket_of(const BucketMap &m, int32_t x) {
    int r = (int)(x / m.owner_block); if (r >= m.R) r = m.R - 1;
    int w = (int)(((int64_t)x - (int64_t)r * m.owner_block) / m.bsize); if (w >= m.nbpr) w = m.nbpr - 1;
    return r * m.nbpr + w;
}

// per own row: #valid edges, #edges into each R_n (cnt is zero elsewhere: memset by the host), records per bucket; act[i] = the row
// has a new-flagged edge (NND_CAND_SKIP; the target side is marked by the import)
__global__ void k_cand_count_sh(const int32_t *ind, const uint8_t *flags, int k, int64_t N, int T, int64_t B, int64_t lo, int64_t hi,
                                uint32_t *cnt, BucketMap bm, unsigned long long *bucket_cnt, uint8_t *act) {
    const int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = i < hi;
    const int lane = threadIdx.x & 31;
    uint32_t valid = 0;
    bool anynew = false;
    if (on)
        for (int n = 0; n < T; ++n) {
            const int64_t bs = (int64_t)n * B, be = bs + B;
            uint32_t c = 0;
            for (int j = 0; j < k; ++j) { int32_t x = ind[i * k + j]; c += (x >= bs && x < be && x >= 0); }
            cnt[(size_t)n * N + i] = c;
        }
    for (int j = 0; j < k; ++j) {
        int o = -1;
        if (on) { const int32_t x = ind[i * k + j]; if (x >= 0) { ++valid; o = bucket_of(bm, x); anynew |= flags[i * k + j] != 0; } }
        const unsigned mk = __match_any_sync(0xffffffffu, o);
        if (o >= 0 && lane == __ffs(mk) - 1) atomicAdd(&bucket_cnt[o], (unsigned long long)__popc(mk));
    }
    if (on) { cnt[(size_t)T * N + i] = valid; if (act && anynew) act[i] = 1; }
}

// one block: exclusive scan of the bucket counts, totals per owner process, cursors cleared
__global__ void k_cand_bucket_scan(unsigned long long *scal, int nbk, int R, int nbpr) {
    __shared__ unsigned long long s[CAND_MAXB];
    const int t = threadIdx.x;
    const unsigned long long mine = t < nbk ? scal[SC_CNT + t] : 0ULL;
    s[t] = mine;
    __syncthreads();
    for (int o = 1; o < CAND_MAXB; o <<= 1) {
        unsigned long long v = t >= o ? s[t - o] : 0ULL;
        __syncthreads();
        s[t] += v;
        __syncthreads();
    }
    if (t < nbk) { scal[SC_OFF + t] = s[t] - mine; scal[SC_CUR + t] = 0ULL; }
    if (t < R) { const int b0 = t * nbpr, b1 = b0 + nbpr - 1; scal[SC_TOT + t] = s[b1] - (b0 > 0 ? s[b0 - 1] : 0ULL); }
    if (t == 0) scal[SC_ALL] = s[nbk - 1];
}

// out[5n..5n+4] = this process's partial {prR_n(s_n), prV(s_n), prV(e_n), prR_n(e_n), total R_n}; out[5T] = total valid
__global__ void k_cand_locals(const uint32_t *pre, int64_t N, int T, int64_t B, long long *out) {
    const int n = threadIdx.x;
    const uint32_t *PV = pre + (size_t)T * N;
    const uint32_t baseV = PV[0], endV = pre[(size_t)(T + 1) * N];
    if (n < T) {
        const uint32_t *PR = pre + (size_t)n * N;
        const uint32_t baseR = PR[0], endR = pre[(size_t)(n + 1) * N];
        const int64_t s = (int64_t)n * B < N ? (int64_t)n * B : N, e = s + B < N ? s + B : N;
        out[5 * n + 0] = (long long)((s < N ? PR[s] : endR) - baseR);
        out[5 * n + 1] = (long long)((s < N ? PV[s] : endV) - baseV);
        out[5 * n + 2] = (long long)((e < N ? PV[e] : endV) - baseV);
        out[5 * n + 3] = (long long)((e < N ? PR[e] : endR) - baseR);
        out[5 * n + 4] = (long long)(endR - baseR);
    }
    if (n == 0) out[5 * T] = (long long)(endV - baseV);
}
// one process = all rows: the global table is the local one, the lower ranks' totals are zero
__global__ void k_cand_glob_single(const long long *loc, int T, uint32_t *glob) {
    const int n = threadIdx.x;
    if (n >= T) return;
    for (int q = 0; q < 4; ++q) glob[6 * n + q] = (uint32_t)loc[5 * n + q];
    glob[6 * n + 4] = 0u; glob[6 * n + 5] = 0u;
}

// glob[6n..6n+5] = {R_s, V_s, V_e, R_e (global), baseR_n, baseV (sums over the lower ranks)}
__device__ __forceinline__ uint32_t draw_index_sh(const uint32_t *pre, const uint32_t *glob, int64_t N, int T, int64_t B, int n, int64_t i,
                                                  uint32_t within_valid, uint32_t within_n) {
    const uint32_t *PR = pre + (size_t)n * N, *PV = pre + (size_t)T * N;
    const uint32_t *G = glob + 6 * n;
    const int64_t s = (int64_t)n * B, e = s + B < N ? s + B : N;
    const uint32_t prR = G[4] + (PR[i] - PR[0]), prV = G[5] + (PV[i] - PV[0]);   // i is an own row: local prefix + lower ranks' totals
    if (i < s) return prR + within_n;
    if (i < e) return G[0] + (prV - G[1]) + within_valid;
    return G[0] + (G[2] - G[1]) + (prR - G[3]) + within_n;
}

// forward priorities of the own rows + one record per reverse edge, appended to the bucket of the target's range
__global__ void k_cand_export(const int32_t *ind, const uint8_t *flags, int k, int64_t N, int T, int64_t B, int64_t lo, int64_t hi,
                              const uint32_t *pre, const uint32_t *glob, const float *table, int64_t table_len, float *fprio,
                              BucketMap bm, const unsigned long long *bucket_off, unsigned long long *bucket_cur, CandRec *out) {
    const int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = i < hi;
    const int lane = threadIdx.x & 31;
    const int ni = on ? (int)(i / B) : 0;
    uint32_t wv = 0;
    for (int j = 0; j < k; ++j) {
        int o = -1; int32_t x = -1; float pr = 0.f;
        if (on) {
            x = ind[i * k + j];
            if (x < 0) fprio[i * k + j] = 0.f;
            else {
                const int nx = (int)(x / B);
                const uint32_t df = draw_index_sh(pre, glob, N, T, B, ni, i, wv, 0);
                const float pf = table[(size_t)ni * table_len + df];
                fprio[i * k + j] = pf;
                pr = pf;
                if (nx != ni) {
                    uint32_t wn = 0;
                    const int64_t bs = (int64_t)nx * B, be = bs + B;
                    for (int jj = 0; jj < j; ++jj) { int32_t y = ind[i * k + jj]; wn += (y >= bs && y < be); }
                    pr = table[(size_t)nx * table_len + draw_index_sh(pre, glob, N, T, B, nx, i, 0, wn)];
                }
                o = bucket_of(bm, x);
                ++wv;
            }
        }
        const unsigned mk = __match_any_sync(0xffffffffu, o);
        const int leader = __ffs(mk) - 1;
        unsigned long long base = 0;
        if (o >= 0 && lane == leader) base = atomicAdd(&bucket_cur[o], (unsigned long long)__popc(mk));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (o >= 0) {
            CandRec r; r.tgt = x; r.key = (uint32_t)(((unsigned long long)(i * k + j) << 1) | (unsigned long long)(flags[i * k + j] != 0)); r.prio = pr;
            out[bucket_off[o] + base + __popc(mk & ((1u << lane) - 1))] = r;
        }
    }
}

// n_dev (optional): the record count lives on the device (single-process path: no host round trip); threads beyond it exit
__global__ void k_cand_import_count(const CandRec *in, unsigned long long n, const unsigned long long *n_dev, int64_t lo, int64_t hi,
                                    uint32_t *indeg, uint8_t *act, unsigned long long *bad) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (n_dev ? *n_dev : n)) return;
    const CandRec r = in[i];
    if (r.tgt < lo || r.tgt >= hi) { atomicAdd(bad, 1ULL); return; }
    atomicAdd(&indeg[r.tgt], 1u);
    if (act && (r.key & 1u)) act[r.tgt] = 1;
}
__global__ void k_cand_import_scatter(const CandRec *in, unsigned long long n, const unsigned long long *n_dev, int64_t lo, int64_t hi,
                                      const uint32_t *revoff, uint32_t *cursor, unsigned long long *rev) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (n_dev ? *n_dev : n)) return;
    const CandRec r = in[i];
    if (r.tgt < lo || r.tgt >= hi) return;
    const uint32_t slot = atomicAdd(&cursor[r.tgt], 1u);
    rev[(size_t)revoff[r.tgt] + slot] = ((unsigned long long)r.key << 32) | (unsigned long long)__float_as_uint(r.prio);
}

static int cand_range(const nnd_ctx *ctx, int64_t N, int64_t *lo, int64_t *hi) {
    *lo = ctx->row_lo > 0 ? (ctx->row_lo < N ? ctx->row_lo : N) : 0;
    *hi = ctx->row_hi < N ? (ctx->row_hi > *lo ? ctx->row_hi : *lo) : N;
    return NND_OK;
}

// step 1 on the device: counts of the own rows, their prefix sums, bucket offsets, this process's count table (W.scal)
static int cand_counts_device(nnd_ctx *ctx, const nnd_problem *P, int T, const BucketMap &bm, CandWs &W, int64_t lo, int64_t hi, uint8_t *act,
                              cudaStream_t st) {
    const int64_t N = P->N; const int k = P->k;
    const int64_t B = N / T + 1;
    NND_CUDA(cudaMemsetAsync(W.cnt, 0, sizeof(uint32_t) * (size_t)(T + 1) * N, st));
    NND_CUDA(cudaMemsetAsync(W.scal, 0, sizeof(unsigned long long) * 4096, st));
    if (act) NND_CUDA(cudaMemsetAsync(act, 0, (size_t)N + 1, st));
    if (hi > lo) { k_cand_count_sh<<<(unsigned)cdiv(hi - lo, 128), 128, 0, st>>>(P->heap_idx, P->heap_flag, k, N, T, B, lo, hi, W.cnt, bm, W.scal + SC_CNT, act); NND_LAUNCH(ctx); }
    NND_TRY(nnd_exclusive_scan_u32(ctx, W.cnt, W.pre, (int64_t)(T + 1) * N, W.scan_tmp, st));
    k_cand_bucket_scan<<<1, CAND_MAXB, 0, st>>>(W.scal, bm.R * bm.nbpr, bm.R, bm.nbpr); NND_LAUNCH(ctx);
    k_cand_locals<<<1, 256, 0, st>>>(W.pre, N, T, B, (long long *)(W.scal + SC_LOC)); NND_LAUNCH(ctx);
    return NND_OK;
}
static int cand_tables(nnd_ctx *ctx, const int64_t *rng_state_host, int T, int tables_valid, CandWs &W, cudaStream_t st) {
    if (tables_valid) return NND_OK;
    for (int n = 0; n < T; ++n) {   // the T priority streams: constant across iterations (utils.py:262 copies rng_state, never advances it)
        int64_t s[3] = {rng_state_host[0] + n, rng_state_host[1] + n, rng_state_host[2] + n};
        NND_TRY(nnd_rng_table_launch(ctx, s, 0, W.table_len, 1, W.table + (size_t)n * W.table_len, st));
    }
    return NND_OK;
}
static int cand_import(nnd_ctx *ctx, const nnd_problem *P, int32_t C, CandWs &W, const CandRec *records, int64_t n_max, const unsigned long long *n_dev,
                       int64_t lo, int64_t hi, uint8_t *act, bool check_bad, int32_t *new_cand, int32_t *old_cand, cudaStream_t st) {
    const int64_t N = P->N;
    NND_CUDA(cudaMemsetAsync(W.indeg, 0, sizeof(uint32_t) * (N + 1), st));
    NND_CUDA(cudaMemsetAsync(W.cursor, 0, sizeof(uint32_t) * (N + 1), st));
    unsigned long long *bad = W.scal + SC_ALL + 1;
    NND_CUDA(cudaMemsetAsync(bad, 0, sizeof(unsigned long long), st));
    if (n_max > 0) { k_cand_import_count<<<(unsigned)cdiv(n_max, 256), 256, 0, st>>>(records, (unsigned long long)n_max, n_dev, lo, hi, W.indeg, act, bad); NND_LAUNCH(ctx); }
    if (check_bad) {
        unsigned long long hbad = 0;
        NND_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(hbad), cudaMemcpyDeviceToHost, st));
        NND_CUDA(cudaStreamSynchronize(st));
        NND_CHECK(hbad == 0, NND_ERR_INVALID, "candidate import: %llu records target rows outside [%lld,%lld)", hbad, (long long)lo, (long long)hi);
    }
    NND_TRY(nnd_exclusive_scan_u32_padded(ctx, W.indeg, W.revoff, N, W.scan_tmp, st));
    if (n_max > 0) { k_cand_import_scatter<<<(unsigned)cdiv(n_max, 256), 256, 0, st>>>(records, (unsigned long long)n_max, n_dev, lo, hi, W.revoff, W.cursor, W.rev); NND_LAUNCH(ctx); }
    NND_STAGE(ctx, NND_ST_CAND_SELECT, st);
    NND_TRY(launch_cand_select(ctx, P, C, W, new_cand, old_cand, act, st));
    return NND_OK;
}

// new_build_candidates on ONE process without a host round trip: counts -> export into range buckets -> import -> selection
int nnd_build_candidates_partitioned(nnd_ctx *ctx, const nnd_problem *P, int32_t C, const int64_t *rng_state_host, int32_t T, int32_t tables_valid,
                                     int32_t *new_cand, int32_t *old_cand, CandWs &W, cudaStream_t st) {
    const int64_t N = P->N; const int k = P->k;
    const int64_t B = N / T + 1;
    const BucketMap bm = bucket_map(N + 1, 1);
    uint8_t *act = ctx->opt_cand_skip ? W.act : nullptr;
    NND_TRY(cand_tables(ctx, rng_state_host, T, tables_valid, W, st));
    NND_TRY(cand_counts_device(ctx, P, T, bm, W, 0, N, act, st));
    uint32_t *dglob = (uint32_t *)(W.scal + SC_GLOB);
    k_cand_glob_single<<<1, 256, 0, st>>>((const long long *)(W.scal + SC_LOC), T, dglob); NND_LAUNCH(ctx);
    k_cand_export<<<(unsigned)cdiv(N, 128), 128, 0, st>>>(P->heap_idx, P->heap_flag, k, N, T, B, 0, N, W.pre, dglob, W.table, W.table_len, W.fprio, bm,
                                                         W.scal + SC_OFF, W.scal + SC_CUR, W.recs);
    NND_LAUNCH(ctx);
    return cand_import(ctx, P, C, W, W.recs, N * (int64_t)k, W.scal + SC_ALL, 0, N, act, false, new_cand, old_cand, st);
}

extern "C" int nnd_cand_local_counts(nnd_ctx *ctx, const nnd_problem *P, int32_t C, int32_t T, int32_t R, int64_t owner_block,
                                     void *ws, size_t ws_bytes, int64_t *out_host, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    NND_CHECK(ctx && P && out_host && T >= 1 && T <= 256 && R >= 1 && R <= 64 && owner_block > 0, NND_ERR_INVALID, "nnd_cand_local_counts: bad args");
    const int64_t N = P->N; const int k = P->k;
    NND_CHECK((double)N * k * 2 < 4.0e9, NND_ERR_UNSUPPORTED, "N*k too large for 32-bit edge keys");
    CandWs W;
    NND_CHECK(carve_cand_ws(ws, ws_bytes, N, k, C, T, &W, nullptr) == NND_OK, NND_ERR_WORKSPACE, "nnd_cand_local_counts: workspace too small");
    int64_t lo, hi; cand_range(ctx, N, &lo, &hi);
    NND_STAGE(ctx, NND_ST_CAND_FILL, st);
    NND_TRY(cand_counts_device(ctx, P, T, bucket_map(owner_block, R), W, lo, hi, ctx->opt_cand_skip ? W.act : nullptr, st));
    NND_CUDA(cudaMemcpyAsync(out_host, W.scal + SC_LOC, sizeof(long long) * (5 * T + 1), cudaMemcpyDeviceToHost, st));
    NND_CUDA(cudaMemcpyAsync(out_host + 5 * T + 1, W.scal + SC_TOT, sizeof(unsigned long long) * R, cudaMemcpyDeviceToHost, st));
    NND_CUDA(cudaStreamSynchronize(st));
    NND_STAGE(ctx, NND_ST_OTHER, st);
    NND_CUDA(cudaGetLastError());
    return NND_OK;
}

extern "C" int nnd_cand_export(nnd_ctx *ctx, const nnd_problem *P, int32_t C, const int64_t *rng_state_host, int32_t T, int32_t tables_valid,
                               int32_t R, int64_t owner_block, const int64_t *glob_host, void *records_out, void *ws, size_t ws_bytes,
                               void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    NND_CHECK(ctx && P && rng_state_host && glob_host && T >= 1 && T <= 256 && R >= 1 && R <= 64 && owner_block > 0, NND_ERR_INVALID,
              "nnd_cand_export: bad args");
    const int64_t N = P->N; const int k = P->k;
    CandWs W;
    NND_CHECK(carve_cand_ws(ws, ws_bytes, N, k, C, T, &W, nullptr) == NND_OK, NND_ERR_WORKSPACE, "nnd_cand_export: workspace too small");
    int64_t lo, hi; cand_range(ctx, N, &lo, &hi);
    const int64_t B = N / T + 1;
    NND_STAGE(ctx, NND_ST_CAND_FILL, st);
    NND_TRY(cand_tables(ctx, rng_state_host, T, tables_valid, W, st));
    std::vector<uint32_t> hg((size_t)6 * T);
    for (int q = 0; q < 6 * T; ++q) { NND_CHECK(glob_host[q] >= 0 && glob_host[q] < (1LL << 32), NND_ERR_INVALID, "nnd_cand_export: count table out of range"); hg[q] = (uint32_t)glob_host[q]; }
    uint32_t *dglob = (uint32_t *)(W.scal + SC_GLOB);
    NND_CUDA(cudaMemcpyAsync(dglob, hg.data(), sizeof(uint32_t) * hg.size(), cudaMemcpyHostToDevice, st));
    NND_CUDA(cudaStreamSynchronize(st));   // the host vector goes out of scope
    if (hi > lo) {
        NND_CHECK(records_out, NND_ERR_INVALID, "nnd_cand_export: records_out is null");
        k_cand_export<<<(unsigned)cdiv(hi - lo, 128), 128, 0, st>>>(P->heap_idx, P->heap_flag, k, N, T, B, lo, hi, W.pre, dglob, W.table, W.table_len,
                                                                W.fprio, bucket_map(owner_block, R), W.scal + SC_OFF, W.scal + SC_CUR,
                                                                (CandRec *)records_out);
        NND_LAUNCH(ctx);
    }
    NND_STAGE(ctx, NND_ST_OTHER, st);
    NND_CUDA(cudaGetLastError());
    return NND_OK;
}

extern "C" int nnd_cand_import_select(nnd_ctx *ctx, const nnd_problem *P, int32_t C, int32_t T, const void *records, int64_t n,
                                      int32_t *new_cand, int32_t *old_cand, void *ws, size_t ws_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    NND_CHECK(ctx && P && new_cand && old_cand && n >= 0 && (records || n == 0), NND_ERR_INVALID, "nnd_cand_import_select: bad args");
    const int64_t N = P->N; const int k = P->k;
    CandWs W;
    NND_CHECK(carve_cand_ws(ws, ws_bytes, N, k, C, T, &W, nullptr) == NND_OK, NND_ERR_WORKSPACE, "nnd_cand_import_select: workspace too small");
    NND_CHECK(n <= 2 * N * (int64_t)k, NND_ERR_OVERFLOW, "nnd_cand_import_select: %lld records exceed the reverse-edge scratch", (long long)n);
    int64_t lo, hi; cand_range(ctx, N, &lo, &hi);
    NND_STAGE(ctx, NND_ST_CAND_FILL, st);
    // W.act still holds the own rows' marks of step 1 (same scratch); the import adds the targets of new-flagged records
    NND_TRY(cand_import(ctx, P, C, W, (const CandRec *)records, n, nullptr, lo, hi, ctx->opt_cand_skip ? W.act : nullptr, true, new_cand, old_cand, st));
    NND_STAGE(ctx, NND_ST_OTHER, st);
    NND_CUDA(cudaGetLastError());
    return NND_OK;
}
