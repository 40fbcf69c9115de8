// k_misc.cu -- context, error text, rng jump tables, scans, heap init / init_random / deheap_sort,
// per-point metric auxiliaries and the nn_descent_internal driver loop.
#include <stdarg.h>

#include <vector>

#include "nnd_common.cuh"

// ------------------------------------------------------------------ errors
static thread_local char g_err[1024] = "";
void nnd_set_error(const char *fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char *nnd_last_error(void) { return g_err; }
extern "C" int nnd_version(void) { return 100; }
extern "C" int64_t nnd_launch_count(nnd_ctx *ctx) { return ctx ? ctx->launches : 0; }

// ------------------------------------------------------------------ context + GF(2) jump tables
static void build_jump_tables(uint64_t *J) {
    // column b of M_c = f_c(e_b); f_c is GF(2)-linear on the 64-bit word (utils.py:30-38: masks,
    // shifts, xor and an arithmetic >> which only copies bits)
    for (int c = 0; c < 3; ++c) {
        uint64_t *M0 = J + ((size_t)c * NND_JUMP_POW) * 64;
        for (int b = 0; b < 64; ++b) {
            long long s[3] = {0, 0, 0};
            s[c] = (long long)(1ULL << b);
            tau_step(s);
            M0[b] = (uint64_t)s[c];
        }
        for (int j = 1; j < NND_JUMP_POW; ++j) {
            const uint64_t *A = J + ((size_t)c * NND_JUMP_POW + (j - 1)) * 64;
            uint64_t *Bm = J + ((size_t)c * NND_JUMP_POW + j) * 64;
            for (int b = 0; b < 64; ++b) {
                uint64_t x = A[b], r = 0;
                for (int t = 0; t < 64; ++t)
                    if ((x >> t) & 1ULL) r ^= A[t];
                Bm[b] = r;
            }
        }
    }
}

extern "C" int nnd_create(nnd_ctx **out, int device) {
    NND_CHECK(out, NND_ERR_INVALID, "nnd_create: null out");
    NND_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    NND_CUDA(cudaGetDeviceProperties(&prop, device));
    NND_CHECK(prop.major >= 10, NND_ERR_UNSUPPORTED, "libnnd_b200 needs an sm_100a device (found sm_%d%d)", prop.major, prop.minor);
    nnd_ctx *c = new nnd_ctx();
    c->device = device; c->sm_count = prop.multiProcessorCount; c->launches = 0;
    memset(c->prof, 0, sizeof(c->prof));
    c->row_lo = 0; c->row_hi = INT64_MAX; c->plan = nullptr; c->plan_free = nullptr; c->jb_ppb_max = 0.0; c->thr_ext = nullptr; c->heap_trusted = 1;
    {
        auto env_int = [](const char *name, int dflt) { const char *e = getenv(name); return (e && *e) ? atoi(e) : dflt; };
        c->opt_rank_scatter = env_int("NND_RANK_SCATTER", 1);
        c->opt_cand_skip = env_int("NND_CAND_SKIP", 1);
        c->opt_cand_partition = env_int("NND_CAND_PARTITION", 0);
        c->opt_leaf_split = env_int("NND_LEAF_SPLIT", 1);
        c->opt_carveout = env_int("NND_SMEM_CARVEOUT", 1);
        c->opt_apply_short = env_int("NND_APPLY_SHORT_SORT", 1);
        c->opt_forest_pm = env_int("NND_FOREST_PM", 1);
        c->opt_forest_pm_mb = env_int("NND_FOREST_PM_MB", 4096);   // 10M: 48/96 MB 250 ms, 200+ MB 241 ms (A/B): point-major on every level
    }
    c->sp_on = 0; c->sp_n = 0; c->sp_cap = 0; c->sp_ev = nullptr; c->sp_id = nullptr; memset(c->sp_ms, 0, sizeof(c->sp_ms));
    { const char *e = getenv("NND_STAGE_PROFILE"); if (e && e[0] == '1') c->sp_on = 1; }
    NND_CUDA(cudaEventCreate(&c->ev0));
    NND_CUDA(cudaEventCreate(&c->ev1));
    size_t n = (size_t)3 * NND_JUMP_POW * 64;
    c->h_jump = (uint64_t *)malloc(n * 8);
    build_jump_tables(c->h_jump);
    NND_CUDA(cudaMalloc(&c->d_jump, n * 8));
    NND_CUDA(cudaMemcpy(c->d_jump, c->h_jump, n * 8, cudaMemcpyHostToDevice));
    *out = c;
    return NND_OK;
}
extern "C" int nnd_destroy(nnd_ctx *ctx) {
    if (!ctx) return NND_OK;
    if (ctx->plan && ctx->plan_free) ctx->plan_free(ctx->plan);
    for (int i = 0; i < ctx->sp_cap; ++i) cudaEventDestroy(ctx->sp_ev[i]);
    free(ctx->sp_ev); free(ctx->sp_id);
    cudaFree(ctx->d_jump);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    free(ctx->h_jump);
    delete ctx;
    return NND_OK;
}
extern "C" int nnd_set_row_range(nnd_ctx *ctx, int64_t lo, int64_t hi) {
    NND_CHECK(ctx && lo >= 0 && hi >= lo, NND_ERR_INVALID, "nnd_set_row_range: bad range");
    ctx->row_lo = lo; ctx->row_hi = hi;
    return NND_OK;
}
// ---- stage profiler: a ring of (event, stage id) marks; the time between two consecutive marks belongs to the earlier mark's stage
static void stage_resolve(nnd_ctx *ctx) {
    if (ctx->sp_n < 2) return;
    cudaEventSynchronize(ctx->sp_ev[ctx->sp_n - 1]);
    for (int i = 0; i + 1 < ctx->sp_n; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->sp_ev[i], ctx->sp_ev[i + 1]) == cudaSuccess) ctx->sp_ms[ctx->sp_id[i]] += ms;
    }
    // the last mark opens an interval that is still running: it becomes the first mark of the next batch
    std::swap(ctx->sp_ev[0], ctx->sp_ev[ctx->sp_n - 1]);
    ctx->sp_id[0] = ctx->sp_id[ctx->sp_n - 1];
    ctx->sp_n = 1;
}
void nnd_stage_mark_slow(nnd_ctx *ctx, int id, cudaStream_t st) {
    if (id < 0 || id >= NND_STAGE_COUNT) id = 0;
    if (ctx->sp_n > 0 && ctx->sp_id[ctx->sp_n - 1] == id) return;      // same stage continues
    if (ctx->sp_cap == 0) {
        ctx->sp_cap = 4096;
        ctx->sp_ev = (cudaEvent_t *)malloc(sizeof(cudaEvent_t) * ctx->sp_cap);
        ctx->sp_id = (int *)malloc(sizeof(int) * ctx->sp_cap);
        for (int i = 0; i < ctx->sp_cap; ++i) cudaEventCreate(&ctx->sp_ev[i]);
    }
    if (ctx->sp_n == ctx->sp_cap) stage_resolve(ctx);
    cudaEventRecord(ctx->sp_ev[ctx->sp_n], st);
    ctx->sp_id[ctx->sp_n] = id;
    ++ctx->sp_n;
}
extern "C" int nnd_stage_profile(nnd_ctx *ctx, int32_t enable) {
    NND_CHECK(ctx, NND_ERR_INVALID, "nnd_stage_profile: null context");
    if (!enable && ctx->sp_on) { stage_resolve(ctx); ctx->sp_n = 0; }
    ctx->sp_on = enable ? 1 : 0;
    return NND_OK;
}
extern "C" int nnd_stage_read(nnd_ctx *ctx, double *out_ms, int32_t reset) {
    NND_CHECK(ctx && out_ms, NND_ERR_INVALID, "nnd_stage_read: bad args");
    stage_resolve(ctx);
    memcpy(out_ms, ctx->sp_ms, sizeof(ctx->sp_ms));
    if (reset) { memset(ctx->sp_ms, 0, sizeof(ctx->sp_ms)); ctx->sp_n = 0; }
    return NND_OK;
}
extern "C" int nnd_set_heap_trusted(nnd_ctx *ctx, int32_t trusted) {
    NND_CHECK(ctx, NND_ERR_INVALID, "nnd_set_heap_trusted: null context");
    ctx->heap_trusted = trusted ? 1 : 0;
    return NND_OK;
}
extern "C" int nnd_set_thresholds(nnd_ctx *ctx, const float *thr) {
    NND_CHECK(ctx, NND_ERR_INVALID, "nnd_set_thresholds: null context");
    ctx->thr_ext = thr;
    return NND_OK;
}
extern "C" int nnd_profile_read(nnd_ctx *ctx, double *out10, int32_t reset) {
    NND_CHECK(ctx && out10, NND_ERR_INVALID, "nnd_profile_read: bad args");
    memcpy(out10, ctx->prof, sizeof(ctx->prof));
    if (reset) memset(ctx->prof, 0, sizeof(ctx->prof));
    return NND_OK;
}
extern "C" int nnd_rng_advance_host(nnd_ctx *ctx, int64_t *state_host, int64_t n) {
    NND_CHECK(ctx && state_host && n >= 0, NND_ERR_INVALID, "nnd_rng_advance_host: bad args");
    long long s[3] = {state_host[0], state_host[1], state_host[2]};
    tau_jump(s, (unsigned long long)n, ctx->h_jump);
    state_host[0] = s[0]; state_host[1] = s[1]; state_host[2] = s[2];
    return NND_OK;
}

// ------------------------------------------------------------------ exclusive scan (u32), 3 kernels
#define SCAN_T 1024
#define SCAN_I 4
#define SCAN_B (SCAN_T * SCAN_I)
size_t nnd_scan_tmp_elems(int64_t n) { return (size_t)cdiv(n, SCAN_B) + 64; }

template <int TF>
__device__ __forceinline__ uint32_t scan_load(const uint32_t *in, int64_t i, int64_t n) {
    if (i >= n) return 0;
    uint32_t v = in[i];
    return TF ? padded_seg(v) : v;
}
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total) {
    __shared__ uint32_t ws[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) ws[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = ws[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
        ws[lane] = wi - w;
        if (lane == 31 && total) *total = wi;
    }
    __syncthreads();
    uint32_t r = ws[warp] + incl - v;
    __syncthreads();
    return r;
}
template <int TF>
__global__ void __launch_bounds__(SCAN_T) k_scan_reduce(const uint32_t *in, int64_t n, uint32_t *bsum) {
    __shared__ uint32_t tot;
    int64_t base = (int64_t)blockIdx.x * SCAN_B + (int64_t)threadIdx.x * SCAN_I;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_I; ++i) s += scan_load<TF>(in, base + i, n);
    block_exclusive_scan(s, &tot);
    if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(SCAN_T) k_scan_bsums(uint32_t *bsum, int64_t nb) {
    __shared__ uint32_t tot;
    uint32_t carry = 0;
    for (int64_t b0 = 0; b0 < nb; b0 += SCAN_T) {
        int64_t i = b0 + threadIdx.x;
        uint32_t v = i < nb ? bsum[i] : 0;
        uint32_t e = block_exclusive_scan(v, &tot);
        if (i < nb) bsum[i] = carry + e;
        carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) bsum[nb] = carry;
}
template <int TF>
__global__ void __launch_bounds__(SCAN_T) k_scan_final(const uint32_t *in, uint32_t *out, int64_t n, const uint32_t *bsum,
                                                        int64_t nb) {
    int64_t base = (int64_t)blockIdx.x * SCAN_B + (int64_t)threadIdx.x * SCAN_I;
    uint32_t v[SCAN_I], s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_I; ++i) { v[i] = scan_load<TF>(in, base + i, n); s += v[i]; }
    uint32_t e = block_exclusive_scan(s, nullptr) + bsum[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_I; ++i) { if (base + i < n) out[base + i] = e; e += v[i]; }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = bsum[nb];
}
template <int TF>
static int scan_impl(nnd_ctx *ctx, const uint32_t *in, uint32_t *out, int64_t n, uint32_t *tmp, cudaStream_t st) {
    if (n <= 0) { NND_CUDA(cudaMemsetAsync(out, 0, sizeof(uint32_t), st)); return NND_OK; }
    int64_t nb = cdiv(n, SCAN_B);
    k_scan_reduce<TF><<<(unsigned)nb, SCAN_T, 0, st>>>(in, n, tmp); NND_LAUNCH(ctx);
    k_scan_bsums<<<1, SCAN_T, 0, st>>>(tmp, nb); NND_LAUNCH(ctx);
    k_scan_final<TF><<<(unsigned)nb, SCAN_T, 0, st>>>(in, out, n, tmp, nb); NND_LAUNCH(ctx);
    NND_CUDA(cudaGetLastError());
    return NND_OK;
}
int nnd_exclusive_scan_u32(nnd_ctx *ctx, const uint32_t *in, uint32_t *out, int64_t n, uint32_t *tmp, cudaStream_t st) {
    return scan_impl<0>(ctx, in, out, n, tmp, st);
}
int nnd_exclusive_scan_u32_padded(nnd_ctx *ctx, const uint32_t *in, uint32_t *out, int64_t n, uint32_t *tmp, cudaStream_t st) {
    return scan_impl<1>(ctx, in, out, n, tmp, st);
}

// ------------------------------------------------------------------ heap init, metric aux
__global__ void k_make_heap(int32_t *idx, float *dist, uint8_t *flag, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { idx[i] = -1; dist[i] = CUDART_INF_F; flag[i] = 0; }
}
extern "C" int nnd_make_heap(nnd_ctx *ctx, int32_t *heap_idx, float *heap_dist, uint8_t *heap_flag, int64_t N, int32_t k,
                             void *stream) {
    NND_CHECK(ctx && heap_idx && heap_dist && heap_flag && N > 0 && k > 0, NND_ERR_INVALID, "nnd_make_heap: bad args");
    ctx->jb_ppb_max = 0.0;   // a new build: the first local join is planned for the worst case
    k_make_heap<<<(unsigned)cdiv(N * k, 256), 256, 0, (cudaStream_t)stream>>>(heap_idx, heap_dist, heap_flag, N * k);
    NND_LAUNCH(ctx);
    NND_CUDA(cudaGetLastError());
    return NND_OK;
}

extern "C" int nnd_metric_aux_bytes(int32_t metric, int64_t N, size_t *bytes) {
    NND_CHECK(bytes && N > 0, NND_ERR_INVALID, "nnd_metric_aux_bytes: bad args");
    if (metric < 0 || metric > NND_METRIC_MAX) {
        nnd_set_error("unknown metric id %d", metric);
        return NND_ERR_UNSUPPORTED;
    }
    const int kind = nnd_metric_aux_kind(metric);
    *bytes = (kind == 1 ? (size_t)N * 4 : kind == 2 ? (size_t)N * 16 : 0) + 16;
    return NND_OK;
}
// distances.py:138 minkowski's p (metric_kwds.values(), pynndescent_.py:1294-1302) -> the first 16 bytes of the aux buffer
extern "C" int nnd_metric_set_param(nnd_ctx *ctx, int32_t metric, void *aux, double value, void *stream) {
    NND_CHECK(ctx && aux, NND_ERR_INVALID, "nnd_metric_set_param: bad args");
    NND_CHECK(metric == NND_METRIC_MINKOWSKI, NND_ERR_UNSUPPORTED, "metric id %d takes no parameter", metric);
    NND_CHECK(value > 0.0 && value < 1.0e6, NND_ERR_INVALID, "minkowski: p = %g is outside (0, 1e6)", value);
    const double h[2] = {value, 1.0 / value};
    NND_CUDA(cudaMemcpyAsync(aux, h, sizeof(h), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    NND_CUDA(cudaStreamSynchronize((cudaStream_t)stream));   // h is a stack array
    return NND_OK;
}
__global__ void k_metric_prepare(const float *X, int64_t N, int d, int ld, int metric, void *aux) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float *x = X + i * ld;
    if (metric == NND_METRIC_ALT_COSINE) {
        float acc = 0.0f;
        for (int e = 0; e < d; ++e) acc = __fmaf_rn(x[e], x[e], acc);
        ((float *)aux)[i] = acc;
    } else if (metric == NND_METRIC_CORRELATION) {
        double s = 0.0;
        for (int e = 0; e < d; ++e) s = __dadd_rn(s, (double)x[e]);
        double mu = __ddiv_rn(s, (double)d), nn = 0.0;
        for (int e = 0; e < d; ++e) { double sx = __dadd_rn((double)x[e], -mu); nn = __fma_rn(sx, sx, nn); }
        ((double *)aux)[2 * i] = mu; ((double *)aux)[2 * i + 1] = nn;
    } else if (nnd_metric_class(metric) == NND_METRIC_ALT_JACCARD) {
        float nz = 0.0f;                               // number of non-zero coordinates (exact small integer)
        for (int e = 0; e < d; ++e) if (x[e] != 0.0f) nz += 1.0f;
        ((float *)aux)[i] = nz;
    } else if (metric == NND_METRIC_ALT_HELLINGER) {
        float s = 0.0f;                                // distances.py:1405 l1 norm (f32 local in the reference)
        for (int e = 0; e < d; ++e) s = __fadd_rn(s, x[e]);
        ((float *)aux)[i] = s;
    }
}
extern "C" int nnd_metric_prepare(nnd_ctx *ctx, const float *X, int64_t N, int32_t d, int32_t ld, int32_t metric, void *aux,
                                  void *stream) {
    NND_CHECK(ctx && X && N > 0 && d > 0 && ld >= d, NND_ERR_INVALID, "nnd_metric_prepare: bad args");
    NND_CHECK(metric >= 0 && metric <= NND_METRIC_MAX, NND_ERR_UNSUPPORTED, "metric id %d not supported on the GPU path", metric);
    if (nnd_metric_aux_kind(metric) == 0) return NND_OK;
    NND_CHECK(aux, NND_ERR_INVALID, "nnd_metric_prepare: aux is null");
    k_metric_prepare<<<(unsigned)cdiv(N, 128), 128, 0, (cudaStream_t)stream>>>(X, N, d, ld, metric, aux);
    NND_LAUNCH(ctx);
    NND_CUDA(cudaGetLastError());
    return NND_OK;
}

__global__ void k_pair_distances(const float *X, int d, int ld, int metric, const void *aux, const int32_t *a,
                                 const int32_t *b, int64_t n, float *out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = pair_distance_seq(X + (int64_t)a[i] * ld, X + (int64_t)b[i] * ld, d, metric, aux, a[i], b[i]);
}
extern "C" int nnd_pair_distances(nnd_ctx *ctx, const nnd_problem *P, const int32_t *a, const int32_t *b, int64_t n,
                                  float *out, void *stream) {
    NND_CHECK(ctx && P && a && b && out && n >= 0, NND_ERR_INVALID, "nnd_pair_distances: bad args");
    if (n == 0) return NND_OK;
    k_pair_distances<<<(unsigned)cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(P->X, P->d, P->ld, P->metric, P->aux, a, b, n, out);
    NND_LAUNCH(ctx);
    NND_CUDA(cudaGetLastError());
    return NND_OK;
}

// ------------------------------------------------------------------ init_random (pynndescent_.py:188-203)
__global__ void k_ir_need(const int32_t *idx, int k, int64_t N, uint32_t *need) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    uint32_t nd = 0;
    if (idx[i * k] < 0) {
        int filled = 0;
        for (int j = 0; j < k; ++j) filled += idx[i * k + j] >= 0;
        nd = (uint32_t)(k - filled);
    }
    need[i] = nd;
}
#define IR_WARPS 4
__global__ void __launch_bounds__(IR_WARPS * 32) k_init_random(const float *X, int64_t N, int d, int ld, int metric,
                                                               const void *aux, int32_t *heap_idx, float *heap_dist,
                                                               uint8_t *heap_flag, int k, const uint32_t *need,
                                                               const uint32_t *need_off, const int32_t *draws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * IR_WARPS + warp;
    if (row >= N) return;
    const uint32_t nd = need[row];
    if (nd == 0) return;
    const size_t stride = ((size_t)k * 9 + 15) & ~size_t(15);
    unsigned char *hp = smem_raw + (size_t)warp * stride;
    float *pri = (float *)hp; int32_t *ind = (int32_t *)(pri + k); uint8_t *flg = (uint8_t *)(ind + k);
    for (int j = lane; j < k; j += 32) { pri[j] = heap_dist[row * k + j]; ind[j] = heap_idx[row * k + j]; flg[j] = heap_flag[row * k + j]; }
    __syncwarp();
    const int32_t *dr = draws + need_off[row];
    for (uint32_t b0 = 0; b0 < nd; b0 += 32) {
        uint32_t t = b0 + lane;
        int32_t cand = -1; float dd = 0.f;
        if (t < nd) {
            int32_t r = dr[t];
            long long a = (r == (int32_t)0x80000000) ? (long long)r : (long long)(r < 0 ? -r : r);  // np.abs on int32 wraps
            long long mm = a % N; if (mm < 0) mm += N;
            cand = (int32_t)mm;
            dd = pair_distance_seq(X + (int64_t)cand * ld, X + row * ld, d, metric, aux, cand, row);
        }
        uint32_t cntb = min(32u, nd - b0);
        for (uint32_t s = 0; s < cntb; ++s) {
            int32_t c = __shfl_sync(0xffffffffu, cand, (int)s);
            float dv = __shfl_sync(0xffffffffu, dd, (int)s);
            warp_heap_push(pri, ind, flg, k, dv, c, (uint8_t)1, lane);
        }
    }
    for (int j = lane; j < k; j += 32) { heap_dist[row * k + j] = pri[j]; heap_idx[row * k + j] = ind[j]; heap_flag[row * k + j] = flg[j]; }
}
int nnd_rng_table_launch(nnd_ctx *ctx, const int64_t *state, int64_t start, int64_t count, int as_float, void *out,
                         cudaStream_t st);

extern "C" int nnd_init_random(nnd_ctx *ctx, const nnd_problem *P, int64_t *rng_state_host, void *ws, size_t ws_bytes,
                               void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    NND_CHECK(ctx && P && rng_state_host && ws, NND_ERR_INVALID, "nnd_init_random: bad args");
    const int64_t N = P->N; const int k = P->k;
    WsAlloc A(ws, ws_bytes);
    uint32_t *need = A.get<uint32_t>(N + 1), *off = A.get<uint32_t>(N + 2);
    uint32_t *tmp = A.get<uint32_t>(nnd_scan_tmp_elems(N + 2));
    NND_CHECK(A.ok, NND_ERR_WORKSPACE, "nnd_init_random: workspace too small");
    NND_STAGE(ctx, NND_ST_INIT_RANDOM, st);
    k_ir_need<<<(unsigned)cdiv(N, 256), 256, 0, st>>>(P->heap_idx, k, N, need); NND_LAUNCH(ctx);
    NND_TRY(nnd_exclusive_scan_u32(ctx, need, off, N, tmp, st));
    uint32_t total = 0;
    NND_CUDA(cudaMemcpyAsync(&total, off + N, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    NND_CUDA(cudaStreamSynchronize(st));
    if (total == 0) return NND_OK;
    int32_t *draws = A.get<int32_t>(total);
    NND_CHECK(A.ok, NND_ERR_WORKSPACE, "nnd_init_random: workspace too small for %u draws", total);
    NND_TRY(nnd_rng_table_launch(ctx, rng_state_host, 0, total, 0, draws, st));
    const size_t stride = ((size_t)k * 9 + 15) & ~size_t(15);
    size_t smem = IR_WARPS * stride;
    NND_CUDA(cudaFuncSetAttribute(k_init_random, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_init_random<<<(unsigned)cdiv(N, IR_WARPS), IR_WARPS * 32, smem, st>>>(P->X, N, P->d, P->ld, P->metric, P->aux, P->heap_idx,
                                                                           P->heap_dist, P->heap_flag, k, need, off, draws);
    NND_LAUNCH(ctx);
    NND_CUDA(cudaGetLastError());
    return nnd_rng_advance_host(ctx, rng_state_host, (int64_t)total);
}

// ------------------------------------------------------------------ deheap_sort (utils.py:189-218) + correction
__device__ __forceinline__ void siftdown_col(float *h1, int32_t *h2, int stride, int size, int elt) {
    while (elt * 2 + 1 < size) {
        int l = elt * 2 + 1, r = l + 1, swap = elt;
        if (h1[swap * stride] < h1[l * stride]) swap = l;
        if (r < size && h1[swap * stride] < h1[r * stride]) swap = r;
        if (swap == elt) break;
        float tf = h1[elt * stride]; h1[elt * stride] = h1[swap * stride]; h1[swap * stride] = tf;
        int32_t ti = h2[elt * stride]; h2[elt * stride] = h2[swap * stride]; h2[swap * stride] = ti;
        elt = swap;
    }
}
__global__ void k_deheap_sort(int32_t *idx, float *dist, int64_t N, int k, int correction, float *out_dist) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = blockDim.x;
    float *sd = (float *)smem_raw;            // [k][T]
    int32_t *si = (int32_t *)(sd + (size_t)k * T);
    const int64_t row0 = (int64_t)blockIdx.x * T;
    // coalesced load of T rows, transposed into smem
    const int64_t nrows = min((int64_t)T, N - row0);
    for (int64_t e = threadIdx.x; e < nrows * k; e += T) {
        int r = (int)(e / k), j = (int)(e % k);
        sd[(size_t)j * T + r] = dist[row0 * k + e];
        si[(size_t)j * T + r] = idx[row0 * k + e];
    }
    __syncthreads();
    if (threadIdx.x < nrows) {
        float *h1 = sd + threadIdx.x; int32_t *h2 = si + threadIdx.x;
        for (int j = k - 1; j > 0; --j) {
            float tf = h1[0]; h1[0] = h1[(size_t)j * T]; h1[(size_t)j * T] = tf;
            int32_t ti = h2[0]; h2[0] = h2[(size_t)j * T]; h2[(size_t)j * T] = ti;
            siftdown_col(h1, h2, T, j, 0);
        }
    }
    __syncthreads();
    for (int64_t e = threadIdx.x; e < nrows * k; e += T) {
        int r = (int)(e / k), j = (int)(e % k);
        float v = sd[(size_t)j * T + r];
        idx[row0 * k + e] = si[(size_t)j * T + r];
        if (out_dist != dist) dist[row0 * k + e] = v;
        out_dist[row0 * k + e] = (correction == NND_CORRECT_SQRT) ? __fsqrt_rn(v) : v;
    }
}
extern "C" int nnd_deheap_sort(nnd_ctx *ctx, int32_t *idx, float *dist, int64_t N, int32_t k, int32_t correction,
                               float *out_dist, void *stream) {
    NND_CHECK(ctx && idx && dist && out_dist && N > 0 && k > 0, NND_ERR_INVALID, "nnd_deheap_sort: bad args");
    NND_CHECK(correction == NND_CORRECT_NONE || correction == NND_CORRECT_SQRT, NND_ERR_UNSUPPORTED,
              "deheap_sort: correction %d is applied on the host (float64 result, distances.py:704-711)", correction);
    int T = 128;
    while (T > 32 && (size_t)T * k * 8 > 96 * 1024) T >>= 1;
    size_t smem = (size_t)T * k * 8;
    NND_CHECK(smem <= 200 * 1024, NND_ERR_UNSUPPORTED, "deheap_sort: k=%d too large", k);
    NND_CUDA(cudaFuncSetAttribute(k_deheap_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NND_STAGE(ctx, NND_ST_DEHEAP, (cudaStream_t)stream);
    k_deheap_sort<<<(unsigned)cdiv(N, T), T, smem, (cudaStream_t)stream>>>(idx, dist, N, k, correction, out_dist);
    NND_LAUNCH(ctx);
    NND_STAGE(ctx, NND_ST_OTHER, (cudaStream_t)stream);
    NND_CUDA(cudaGetLastError());
    return NND_OK;
}
