// nnd_common.cuh -- shared device/host helpers of libnnd_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/nnd_b200.h"

// ------------------------------------------------------------------ errors / context
void nnd_set_error(const char *fmt, ...);

#define NND_CUDA(call)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (call);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            nnd_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return NND_ERR_CUDA;                                                                         \
        }                                                                                                \
    } while (0)
#define NND_CHECK(cond, code, ...)  \
    do {                            \
        if (!(cond)) {              \
            nnd_set_error(__VA_ARGS__); \
            return (code);          \
        }                           \
    } while (0)
#define NND_TRY(call)              \
    do {                           \
        int _r = (call);           \
        if (_r != NND_OK) return _r; \
    } while (0)

struct nnd_ctx {
    int device;
    int sm_count;
    int64_t launches;
    uint64_t *d_jump;  // [3][NND_JUMP_POW][64] column-major GF(2) matrices M^(2^j)
    uint64_t *h_jump;
    // profile accumulators per join mode (0 local join, 1 leaf): device ms, launches, rows, pairs, pushes
    double prof[2][5];
    cudaEvent_t ev0, ev1;
    unsigned long long last_round_rows, last_round_pairs;
    double jb_ppb_max;        // half-updates per batch of the fullest round since the last plan (0 = unknown): sizes the next plan's rounds
    int heap_trusted;         // nnd_set_heap_trusted: every distance in the heap is the metric's own value (enables the presence / duplicate filters)
    const float *thr_ext;     // nnd_set_thresholds: heap-root snapshot of all rows supplied by the caller (sharded build), or null
    int64_t row_lo, row_hi;   // sharded build: this process joins / samples candidates only for rows in [row_lo,row_hi)
    void *plan;               // JoinPlan of the sharded round API (k_descent.cu)
    void (*plan_free)(void *);
    // ---- stage profiler (nnd_stage_profile / nnd_stage_read): events recorded at stage boundaries, resolved lazily
    int sp_on, sp_n, sp_cap;
    cudaEvent_t *sp_ev; int *sp_id;
    double sp_ms[NND_STAGE_COUNT];
    // ---- schedule options (read from the environment by nnd_create; none of them changes any result)
    int opt_apply_short;      // NND_APPLY_SHORT_SORT=1: k_apply sorts only next_pow2(kept) lanes
    int opt_carveout;         // NND_SMEM_CARVEOUT=1: prefer the maximum shared-memory carveout for the occupancy-by-smem kernels
    int opt_leaf_split;       // NND_LEAF_SPLIT=1: leaf init = distance-matrix kernel + thread-per-member push kernel (k_descent.cu)
    int opt_cand_partition;   // NND_CAND_PARTITION=1: one process routes its reverse edges through range buckets like the sharded build does
                              // (default 0: the direct scatter is faster on one GPU -- cand_fill 120 vs 196 ms at 10M, profiles/ab_r02_10m_call6.json)
    int opt_cand_skip;        // NND_CAND_SKIP=1: candidates are sampled only for vertices that will have NEW candidates (k_cand.cu)
    int opt_forest_pm;        // NND_FOREST_PM=1: point-major hyperplane pass of the RP forest (k_forest_side_pm<NCH>, k_forest.cu)
    int opt_forest_pm_mb;     // NND_FOREST_PM_MB: ... while the level's hyperplane table is at most this many MB (default 48)
    int opt_rank_scatter;     // NND_RANK_SCATTER=1: joins count targets with fire-and-forget reductions, k_scatter_rank ranks
};
// "from here on the stream's time belongs to stage `id`" (no-op unless profiling is on)
void nnd_stage_mark_slow(nnd_ctx *ctx, int id, cudaStream_t st);
#define NND_STAGE(ctx, id, st) do { if ((ctx)->sp_on) nnd_stage_mark_slow((ctx), (id), (st)); } while (0)
#define NND_JUMP_POW 44
#define NND_LAUNCH(ctx) ((ctx)->launches++)

// bump allocator over the caller's workspace
struct WsAlloc {
    char *base;
    size_t cap, off;
    bool ok;
    __host__ WsAlloc(void *p, size_t c) : base((char *)p), cap(c), off(0), ok(true) {}
    template <typename T>
    __host__ T *get(size_t n) {
        size_t bytes = (n * sizeof(T) + 255) & ~size_t(255);
        if (base == nullptr) { off += bytes; return nullptr; }  // sizing pass
        if (off + bytes > cap) { ok = false; off += bytes; return nullptr; }
        T *r = (T *)(base + off);
        off += bytes;
        return r;
    }
};

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------ canonical math
// deterministic log2 in double: only + - * / fma, so CPU oracle and GPU agree bit-for-bit
__device__ __forceinline__ double det_log2(double x) {
    long long u = __double_as_longlong(x);
    int e = (int)((u >> 52) & 0x7FF) - 1023;
    u = (u & 0x000FFFFFFFFFFFFFLL) | 0x3FF0000000000000LL;
    double m = __longlong_as_double(u);
    if (m > 1.4142135623730951) { m = __dmul_rn(m, 0.5); e += 1; }
    double t = __ddiv_rn(__dadd_rn(m, -1.0), __dadd_rn(m, 1.0));
    double t2 = __dmul_rn(t, t);
    double p = 1.0 / 27.0;
    p = __fma_rn(p, t2, 1.0 / 25.0);
    p = __fma_rn(p, t2, 1.0 / 23.0);
    p = __fma_rn(p, t2, 1.0 / 21.0);
    p = __fma_rn(p, t2, 1.0 / 19.0);
    p = __fma_rn(p, t2, 1.0 / 17.0);
    p = __fma_rn(p, t2, 1.0 / 15.0);
    p = __fma_rn(p, t2, 1.0 / 13.0);
    p = __fma_rn(p, t2, 1.0 / 11.0);
    p = __fma_rn(p, t2, 1.0 / 9.0);
    p = __fma_rn(p, t2, 1.0 / 7.0);
    p = __fma_rn(p, t2, 1.0 / 5.0);
    p = __fma_rn(p, t2, 1.0 / 3.0);
    p = __fma_rn(p, t2, 1.0);
    double ln_m = __dmul_rn(__dmul_rn(2.0, t), p);
    return __fma_rn(ln_m, 1.4426950408889634, (double)e);
}

// deterministic 2^z in double (same contract as det_log2; oracle/nnd_oracle.c det_exp2): z = n + f, |f| <= 1/2,
// 2^f = exp(f ln 2) by a degree-13 Horner chain, scaled by 2^n through the exponent bits
__device__ __forceinline__ double det_exp2(double z) {
    if (z > 1023.0) return CUDART_INF;
    if (z < -1021.0) return 0.0;
    const double n = rint(z);
    const double g = __dmul_rn(__dadd_rn(z, -n), 0.6931471805599453);
    double q = 1.0 / 6227020800.0;
    q = __fma_rn(q, g, 1.0 / 479001600.0);
    q = __fma_rn(q, g, 1.0 / 39916800.0);
    q = __fma_rn(q, g, 1.0 / 3628800.0);
    q = __fma_rn(q, g, 1.0 / 362880.0);
    q = __fma_rn(q, g, 1.0 / 40320.0);
    q = __fma_rn(q, g, 1.0 / 5040.0);
    q = __fma_rn(q, g, 1.0 / 720.0);
    q = __fma_rn(q, g, 1.0 / 120.0);
    q = __fma_rn(q, g, 1.0 / 24.0);
    q = __fma_rn(q, g, 1.0 / 6.0);
    q = __fma_rn(q, g, 0.5);
    q = __fma_rn(q, g, 1.0);
    q = __fma_rn(q, g, 1.0);
    return __dmul_rn(q, __longlong_as_double(((long long)n + 1023LL) << 52));
}
// x^y for x >= 0, y > 0 (minkowski): 0^y = 0 -- so zero-padded columns contribute nothing -- else 2^(y log2 x)
__device__ __forceinline__ double det_pow(double x, double y) { return x > 0.0 ? det_exp2(__dmul_rn(y, det_log2(x))) : 0.0; }

#define NND_FLT_MAX 3.402823466e+38f

// distances.py:583-630 tail of alternative_cosine given dot and the two squared norms
__device__ __forceinline__ float finish_alt_cosine(float dot, float nx, float ny) {
    if (nx == 0.0f && ny == 0.0f) return 0.0f;
    if (nx == 0.0f || ny == 0.0f) return NND_FLT_MAX;
    if (dot <= 0.0f) return NND_FLT_MAX;
    float r = __fdiv_rn(__fsqrt_rn(__fmul_rn(nx, ny)), dot);
    return (float)det_log2((double)r);
}
// distances.py:665-701 tail of alternative_dot
// ... and of alternative_inner_product (distances.py:759-790; same class, `metric` picks the tail): 1.0 / result with an f32
// `result` -- the reference's f64 quotient rounded to f32 equals the f32 quotient (53 >= 2*24+2 bits: double rounding is innocuous)
__device__ __forceinline__ float finish_alt_dot(float dot, int metric) {
    if (dot <= 0.0f) return NND_FLT_MAX;
    if (metric == NND_METRIC_ALT_INNER) return __fdiv_rn(1.0f, dot);
    return (float)(-det_log2((double)dot));
}
// distances.py:1283-1320 tail of correlation given the centred dot and squared norms (f64)
__device__ __forceinline__ float finish_correlation(double dp, double nx, double ny) {
    if (nx == 0.0 && ny == 0.0) return 0.0f;
    if (dp == 0.0) return 1.0f;
    return (float)__dadd_rn(1.0, -__ddiv_rn(dp, __dsqrt_rn(__dmul_rn(nx, ny))));
}

// ---- SURVEY 8f-2 metrics ---------------------------------------------------------------------------------
// metric class = which per-element accumulation a kernel instantiation runs; the binary family shares one
__host__ __device__ __forceinline__ int nnd_metric_class(int metric) {
    if (metric == NND_METRIC_ALT_INNER) return NND_METRIC_ALT_DOT;      // same dot accumulation, another tail
    return (metric >= NND_METRIC_ALT_JACCARD && metric <= NND_METRIC_YULE) ? NND_METRIC_ALT_JACCARD : metric;
}
// classes whose accumulator is f64 (the reference accumulates these in a float64 `result`)
__host__ __device__ __forceinline__ bool nnd_metric_acc_f64(int metric) {
    return metric == NND_METRIC_CORRELATION || metric == NND_METRIC_MANHATTAN || metric == NND_METRIC_CANBERRA || metric == NND_METRIC_MINKOWSKI;
}
// per-row auxiliary: 0 none, 1 one f32 per row (cosine: squared norm; binary family: non-zero count;
// hellinger: l1 norm), 2 two f64 per row (correlation: mean, centred squared norm)
__host__ __device__ __forceinline__ int nnd_metric_aux_kind(int metric) {
    const int c = nnd_metric_class(metric);
    if (c == NND_METRIC_ALT_COSINE || c == NND_METRIC_ALT_JACCARD || c == NND_METRIC_ALT_HELLINGER) return 1;
    if (c == NND_METRIC_CORRELATION) return 2;
    return 0;
}
// one element of the accumulation (template class MC); padded columns (both 0) contribute nothing
template <int MC, typename ACC>
__device__ __forceinline__ ACC nnd_metric_step(ACC acc, float xa, float xb) {
    if (MC == NND_METRIC_MANHATTAN) return (ACC)__dadd_rn((double)acc, (double)fabsf(__fsub_rn(xa, xb)));
    if (MC == NND_METRIC_CHEBYSHEV) return (ACC)fmaxf((float)acc, fabsf(__fsub_rn(xa, xb)));
    if (MC == NND_METRIC_CANBERRA) {
        const float den = __fadd_rn(fabsf(xa), fabsf(xb));
        return den > 0.0f ? (ACC)__dadd_rn((double)acc, (double)__fdiv_rn(fabsf(__fsub_rn(xa, xb)), den)) : acc;
    }
    if (MC == NND_METRIC_HAMMING) return (xa != xb) ? (ACC)__fadd_rn((float)acc, 1.0f) : acc;
    if (MC == NND_METRIC_ALT_JACCARD) return (xa != 0.0f && xb != 0.0f) ? (ACC)__fadd_rn((float)acc, 1.0f) : acc;
    if (MC == NND_METRIC_ALT_HELLINGER) return (ACC)__fadd_rn((float)acc, __fsqrt_rn(__fmul_rn(xa, xb)));
    return acc;
}
// distances.py:138-153 minkowski: one element of the sum (p = first double of the metric aux buffer) and the tail
__device__ __forceinline__ double nnd_minkowski_step(double acc, float xa, float xb, double p) {
    return __dadd_rn(acc, det_pow((double)fabsf(__fsub_rn(xa, xb)), p));
}
__device__ __forceinline__ float finish_minkowski(double acc, double inv_p) { return (float)det_pow(acc, inv_p); }
// tails of the binary family from tt = #{both non-zero}, the two non-zero counts and the dimension (all exact
// small integers; the reference does the final division in float64)
__device__ __forceinline__ float finish_binary(int metric, float ttf, float nxf, float nyf, int dim) {
    const double tt = (double)ttf, nx = (double)nxf, ny = (double)nyf, n = (double)dim;
    const double neq = nx + ny - 2.0 * tt;      // #{exactly one non-zero}
    switch (metric) {
    case NND_METRIC_ALT_JACCARD: {              // distances.py:303-326
        const double nnz = nx + ny - tt;
        if (nnz == 0.0) return 0.0f;
        if (tt == 0.0) return CUDART_INF_F;     // -log2(0)
        return (float)(-det_log2((double)__fdiv_rn(ttf, (float)nnz)));   // the reference divides in f32 (locals)
    }
    case NND_METRIC_DICE: return neq == 0.0 ? 0.0f : (float)__ddiv_rn(neq, 2.0 * tt + neq);
    case NND_METRIC_MATCHING: return (float)__ddiv_rn(neq, n);
    case NND_METRIC_KULSINSKI: return neq == 0.0 ? 0.0f : (float)__ddiv_rn(neq - tt + n, neq + n);
    case NND_METRIC_ROGERSTANIMOTO: return (float)__ddiv_rn(2.0 * neq, n + neq);
    case NND_METRIC_RUSSELLRAO: return (tt == nx && tt == ny) ? 0.0f : (float)__ddiv_rn(n - tt, n);
    case NND_METRIC_SOKALSNEATH: return neq == 0.0 ? 0.0f : (float)__ddiv_rn(neq, 0.5 * tt + neq);
    default: {                                  // NND_METRIC_YULE
        const double tf = nx - tt, ft = ny - tt, ff = n - tt - tf - ft;
        if (tf == 0.0 || ft == 0.0) return 0.0f;
        return (float)__ddiv_rn(2.0 * tf * ft, tt * ff + tf * ft);
    }
    }
}
// distances.py:1387-1416 tail of alternative_hellinger given s = sum sqrt(x_i y_i) and the two l1 norms: all
// f32 in the reference (numba locals) and literally the tail of alternative_cosine with (dot, nx, ny) = (s, lx, ly)
__device__ __forceinline__ float finish_alt_hellinger(float s, float lx, float ly) { return finish_alt_cosine(s, lx, ly); }
// tail for the classes without per-row auxiliaries
__device__ __forceinline__ float finish_plain(int metric, double acc, int dim) {
    if (metric == NND_METRIC_HAMMING) return (float)__ddiv_rn(acc, (double)dim);
    return (float)acc;   // manhattan, canberra (f64 -> f32), chebyshev
}

// one whole pair distance by ONE thread straight from global memory (canonical sequential
// chain).  Used by the low-volume paths (init_random, tests); the join kernel has its own
// shared-memory tiled loop with the same per-pair operation order.
__device__ __forceinline__ float pair_distance_seq(const float *__restrict__ x, const float *__restrict__ y, int d,
                                                   int metric, const void *aux, int64_t ix, int64_t iy) {
    if (metric == NND_METRIC_SQEUCLIDEAN) {
        float acc = 0.0f;
        for (int e = 0; e < d; ++e) { float df = __fsub_rn(x[e], y[e]); acc = __fmaf_rn(df, df, acc); }
        return acc;
    } else if (metric == NND_METRIC_ALT_COSINE || metric == NND_METRIC_ALT_DOT || metric == NND_METRIC_ALT_INNER) {
        float acc = 0.0f;
        for (int e = 0; e < d; ++e) acc = __fmaf_rn(x[e], y[e], acc);
        if (metric != NND_METRIC_ALT_COSINE) return finish_alt_dot(acc, metric);
        const float *n2 = (const float *)aux;
        return finish_alt_cosine(acc, n2[ix], n2[iy]);
    } else if (metric == NND_METRIC_CORRELATION) {
        const double *mu = (const double *)aux;
        double mx = mu[2 * ix], my = mu[2 * iy], dp = 0.0;
        for (int e = 0; e < d; ++e) {
            double sx = __dadd_rn((double)x[e], -mx), sy = __dadd_rn((double)y[e], -my);
            dp = __fma_rn(sx, sy, dp);
        }
        return finish_correlation(dp, mu[2 * ix + 1], mu[2 * iy + 1]);
    }
    const int mc = nnd_metric_class(metric);
    if (mc == NND_METRIC_MINKOWSKI) {
        const double p = ((const double *)aux)[0], inv_p = ((const double *)aux)[1];
        double acc = 0.0;
        for (int e = 0; e < d; ++e) acc = nnd_minkowski_step(acc, x[e], y[e], p);
        return finish_minkowski(acc, inv_p);
    }
    if (mc == NND_METRIC_MANHATTAN) {
        double acc = 0.0;
        for (int e = 0; e < d; ++e) acc = nnd_metric_step<NND_METRIC_MANHATTAN, double>(acc, x[e], y[e]);
        return finish_plain(metric, acc, d);
    } else if (mc == NND_METRIC_CHEBYSHEV) {
        float acc = 0.0f;
        for (int e = 0; e < d; ++e) acc = nnd_metric_step<NND_METRIC_CHEBYSHEV, float>(acc, x[e], y[e]);
        return acc;
    } else if (mc == NND_METRIC_CANBERRA) {
        double acc = 0.0;
        for (int e = 0; e < d; ++e) acc = nnd_metric_step<NND_METRIC_CANBERRA, double>(acc, x[e], y[e]);
        return finish_plain(metric, acc, d);
    } else if (mc == NND_METRIC_HAMMING) {
        float acc = 0.0f;
        for (int e = 0; e < d; ++e) acc = nnd_metric_step<NND_METRIC_HAMMING, float>(acc, x[e], y[e]);
        return finish_plain(metric, (double)acc, d);
    } else if (mc == NND_METRIC_ALT_JACCARD) {
        float acc = 0.0f;
        for (int e = 0; e < d; ++e) acc = nnd_metric_step<NND_METRIC_ALT_JACCARD, float>(acc, x[e], y[e]);
        const float *nz = (const float *)aux;
        return finish_binary(metric, acc, nz[ix], nz[iy], d);
    } else {
        float acc = 0.0f;
        for (int e = 0; e < d; ++e) acc = nnd_metric_step<NND_METRIC_ALT_HELLINGER, float>(acc, x[e], y[e]);
        const float *l1 = (const float *)aux;
        return finish_alt_hellinger(acc, l1[ix], l1[iy]);
    }
}

// ------------------------------------------------------------------ Tausworthe rng
// utils.py:17-40; int64 state words, arithmetic >>, not confined to 32 bits
__host__ __device__ __forceinline__ int32_t tau_step(long long *s) {
    s[0] = (long long)((((unsigned long long)(s[0] & 4294967294LL) << 12) & 0xFFFFFFFFULL)) ^
           ((long long)((((unsigned long long)s[0] << 13) & 0xFFFFFFFFULL) ^ (unsigned long long)s[0]) >> 19);
    s[1] = (long long)((((unsigned long long)(s[1] & 4294967288LL) << 4) & 0xFFFFFFFFULL)) ^
           ((long long)((((unsigned long long)s[1] << 2) & 0xFFFFFFFFULL) ^ (unsigned long long)s[1]) >> 25);
    s[2] = (long long)((((unsigned long long)(s[2] & 4294967280LL) << 17) & 0xFFFFFFFFULL)) ^
           ((long long)((((unsigned long long)s[2] << 3) & 0xFFFFFFFFULL) ^ (unsigned long long)s[2]) >> 11);
    return (int32_t)(uint32_t)((unsigned long long)(s[0] ^ s[1] ^ s[2]) & 0xFFFFFFFFULL);
}
// utils.py:43-57 tau_rand
__host__ __device__ __forceinline__ float tau_to_float(int32_t r) {
    double v = (double)r / 2147483647.0;
    return (float)(v < 0 ? -v : v);
}
// x <- M^(2^j) x for every set bit j of n ; jump = [3][NND_JUMP_POW][64]
__host__ __device__ inline void tau_jump(long long *s, unsigned long long n, const uint64_t *jump) {
    for (int j = 0; n != 0 && j < NND_JUMP_POW; ++j, n >>= 1) {
        if (!(n & 1)) continue;
        for (int c = 0; c < 3; ++c) {
            const uint64_t *M = jump + ((size_t)c * NND_JUMP_POW + j) * 64;
            unsigned long long x = (unsigned long long)s[c], r = 0;
            for (int b = 0; b < 64; ++b)
                if ((x >> b) & 1ULL) r ^= M[b];
            s[c] = (long long)r;
        }
    }
}

// 64-bit mixer used for the level-forest node randomness (oracle: mix64/node_key)
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// ------------------------------------------------------------------ heap push (smem rows)
// utils.py:471-533 checked_flagged_heap_push executed by a whole warp on a row held in
// shared memory: duplicate test is lane-parallel, the sift-down is done by lane 0.
// Returns 1 if the push modified the row (uniform across the warp).
__device__ __forceinline__ int warp_heap_push(float *pri, int32_t *ind, uint8_t *flg, int size, float p, int32_t n,
                                              uint8_t f, int lane) {
    if (p >= pri[0]) return 0;
    bool dup = false;
    for (int i = lane; i < size; i += 32) dup |= (ind[i] == n);
    if (__any_sync(0xffffffffu, dup)) return 0;
    if (lane == 0) {
        int i = 0, i_swap;
        for (;;) {
            int ic1 = 2 * i + 1, ic2 = ic1 + 1;
            if (ic1 >= size) break;
            else if (ic2 >= size) { if (pri[ic1] > p) i_swap = ic1; else break; }
            else if (pri[ic1] >= pri[ic2]) { if (p < pri[ic1]) i_swap = ic1; else break; }
            else { if (p < pri[ic2]) i_swap = ic2; else break; }
            pri[i] = pri[i_swap]; ind[i] = ind[i_swap];
            if (flg) flg[i] = flg[i_swap];
            i = i_swap;
        }
        pri[i] = p; ind[i] = n;
        if (flg) flg[i] = f;
    }
    __syncwarp();
    return 1;
}

// ------------------------------------------------------------------ warp bitonic sort (u64)
// n must be a power of two; works on shared or global memory; ascending.
__device__ __forceinline__ void warp_bitonic_sort_u64(unsigned long long *a, int n, int lane) {
    for (int k2 = 2; k2 <= n; k2 <<= 1) {
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < n; i += 32) {
                int l = i ^ j;
                if (l > i) {
                    unsigned long long x = a[i], y = a[l];
                    bool up = ((i & k2) == 0);
                    if ((x > y) == up) { a[i] = y; a[l] = x; }
                }
            }
            __syncwarp();
        }
    }
}
// same, sorting 16-byte records by their first 8 bytes (key), payload carried along
struct __align__(16) Rec16 { unsigned long long key; int32_t oth; float d; };
__device__ __forceinline__ void warp_bitonic_sort_rec(Rec16 *a, int n, int lane) {
    for (int k2 = 2; k2 <= n; k2 <<= 1) {
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < n; i += 32) {
                int l = i ^ j;
                if (l > i) {
                    Rec16 x = a[i], y = a[l];
                    bool up = ((i & k2) == 0);
                    if ((x.key > y.key) == up) { a[i] = y; a[l] = x; }
                }
            }
            __syncwarp();
        }
    }
}
__host__ __device__ __forceinline__ uint32_t next_pow2_u32(uint32_t v) {
    if (v <= 1) return 1;
    v--; v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
    return v + 1;
}
// segments longer than this are padded to a power of two so they can be sorted in place
#define NND_SEG_SMEM 128
__host__ __device__ __forceinline__ uint32_t padded_seg(uint32_t n) { return n <= NND_SEG_SMEM ? n : next_pow2_u32(n); }

// Round a pointer into the dynamic shared array up to `a` bytes (a power of two <= the array's declared alignment) WITHOUT an
// integer round trip: `(T *)(((uintptr_t)p + a - 1) & ~(a - 1))` is an inttoptr, which drops the shared address space -- every
// access through it becomes a generic LD.E / ST.E with 64-bit address arithmetic (and a long-scoreboard dependency) instead of
// LDS / STS with an immediate offset.  Same address as the absolute rounding because `base` is aligned to >= a.
__device__ __forceinline__ unsigned char *smem_align_up(unsigned char *base, unsigned char *p, unsigned a) {
    const unsigned off = (unsigned)(p - base);
    return base + ((off + (a - 1u)) & ~(a - 1u));
}

// ------------------------------------------------------------------ PTX: mbarrier + bulk copy (TMA engine, 1-D)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// cp.async.bulk global -> shared::cta, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------ host-visible kernels' launch helpers
int nnd_exclusive_scan_u32(nnd_ctx *ctx, const uint32_t *in, uint32_t *out, int64_t n, uint32_t *block_tmp,
                           cudaStream_t st);  // out has n+1 entries (out[n] = total)
int nnd_exclusive_scan_u32_padded(nnd_ctx *ctx, const uint32_t *in, uint32_t *out, int64_t n, uint32_t *block_tmp,
                                  cudaStream_t st);  // scans padded_seg(in[i])
size_t nnd_scan_tmp_elems(int64_t n);
int nnd_rng_table_launch(nnd_ctx *ctx, const int64_t *state, int64_t start, int64_t count, int as_float, void *out,
                         cudaStream_t st);
size_t nnd_cand_ws_bytes(int64_t N, int k, int C, int T);
