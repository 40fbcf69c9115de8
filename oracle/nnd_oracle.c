/*
 * nnd_oracle.c -- CPU ORACLE for the NN-descent hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is a plain-C restatement of the reference's numba algorithm
 * (lmcinnes/pynndescent v0.6.0).  It exists so that tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs have something to check the
 * CUDA path against.  Nothing under pynndescent_b200/ may import, link or call it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference_golden.py checks this file
 * against vectors generated from the real reference (tests/golden/make_golden.py,
 * run in the build container where /root/reference is importable).
 *
 * Each function cites the reference file:line it follows.  The ONLY deliberate
 * deviation is floating-point summation ORDER: the reference compiles with numba
 * fastmath (LLVM re-associates/vectorises sums, so its low bits depend on host ISA);
 * here every reduction has one defined order ("canonical order")
 * which the CUDA kernels reproduce bit-for-bit:
 *   - pair distances: one sequential chain  acc = fma(a,b,acc)  over e = 0..d-1
 *   - RP-tree margins: 32 interleaved partial sums (float4 granularity) + xor butterfly
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* metric ids (restated; the product's copy lives in include/nnd_b200.h) */
enum { ORC_SQEUCLIDEAN = 0, ORC_ALT_COSINE = 1, ORC_ALT_DOT = 2, ORC_CORRELATION = 3,
       /* SURVEY 8f-2 */
       ORC_MANHATTAN = 4, ORC_CHEBYSHEV = 5, ORC_CANBERRA = 6, ORC_HAMMING = 7, ORC_ALT_JACCARD = 8, ORC_DICE = 9,
       ORC_MATCHING = 10, ORC_KULSINSKI = 11, ORC_ROGERSTANIMOTO = 12, ORC_RUSSELLRAO = 13, ORC_SOKALSNEATH = 14,
       ORC_YULE = 15, ORC_ALT_HELLINGER = 16, ORC_ALT_INNER = 17, ORC_MINKOWSKI = 18 };

/* ------------------------------------------------------------------ RNG ---- */
/* utils.py:17-40 tau_rand_int.  State words are int64 and are NOT confined to 32
 * bits (the "^ s" term is unmasked, shifts are arithmetic).                      */
ORC_API int32_t orc_tau_rand_int(int64_t *s) {
    s[0] = (int64_t)((((uint64_t)(s[0] & 4294967294LL) << 12) & 0xFFFFFFFFULL)) ^
           ((int64_t)((((uint64_t)s[0] << 13) & 0xFFFFFFFFULL) ^ (uint64_t)s[0]) >> 19);
    s[1] = (int64_t)((((uint64_t)(s[1] & 4294967288LL) << 4) & 0xFFFFFFFFULL)) ^
           ((int64_t)((((uint64_t)s[1] << 2) & 0xFFFFFFFFULL) ^ (uint64_t)s[1]) >> 25);
    s[2] = (int64_t)((((uint64_t)(s[2] & 4294967280LL) << 17) & 0xFFFFFFFFULL)) ^
           ((int64_t)((((uint64_t)s[2] << 3) & 0xFFFFFFFFULL) ^ (uint64_t)s[2]) >> 11);
    return (int32_t)(uint32_t)((uint64_t)(s[0] ^ s[1] ^ s[2]) & 0xFFFFFFFFULL);
}

/* utils.py:43-57 tau_rand: abs(float(int)/0x7FFFFFFF) returned as f4 */
ORC_API float orc_tau_rand(int64_t *s) {
    int32_t r = orc_tau_rand_int(s);
    return (float)fabs((double)r / 2147483647.0);
}

static inline int64_t floormod(int64_t a, int64_t n) {
    int64_t m = a % n;
    if (m < 0) m += n;
    return m;
}
/* np.abs on an int32 wraps for INT32_MIN */
static inline int64_t abs_i32(int32_t r) { return (r == INT32_MIN) ? (int64_t)INT32_MIN : (int64_t)(r < 0 ? -r : r); }

/* ------------------------------------------------------------ distances ---- */
/* deterministic log2 in double using only + - * / fma (bit-identical on CPU and GPU) */
static double det_log2(double x) {
    union { double d; uint64_t u; } v;
    v.d = x;
    int e = (int)((v.u >> 52) & 0x7FF) - 1023;
    v.u = (v.u & 0x000FFFFFFFFFFFFFULL) | 0x3FF0000000000000ULL; /* m in [1,2) */
    double m = v.d;
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    double t = (m - 1.0) / (m + 1.0);
    double t2 = t * t;
    double p = 1.0 / 27.0;
    p = fma(p, t2, 1.0 / 25.0);
    p = fma(p, t2, 1.0 / 23.0);
    p = fma(p, t2, 1.0 / 21.0);
    p = fma(p, t2, 1.0 / 19.0);
    p = fma(p, t2, 1.0 / 17.0);
    p = fma(p, t2, 1.0 / 15.0);
    p = fma(p, t2, 1.0 / 13.0);
    p = fma(p, t2, 1.0 / 11.0);
    p = fma(p, t2, 1.0 / 9.0);
    p = fma(p, t2, 1.0 / 7.0);
    p = fma(p, t2, 1.0 / 5.0);
    p = fma(p, t2, 1.0 / 3.0);
    p = fma(p, t2, 1.0);
    double ln_m = 2.0 * t * p;
    return fma(ln_m, 1.4426950408889634, (double)e);
}
ORC_API double orc_det_log2(double x) { return det_log2(x); }
/* deterministic 2^z in double (same contract): z = n + f, |f| <= 1/2, 2^f = exp(f ln 2) by a degree-13 Horner chain of fma
 * (|f ln 2| <= 0.3466: the first dropped term is 0.3466^14 / 14! = 4e-18), scaled by 2^n through the exponent bits */
static double det_exp2(double z) {
    if (z > 1023.0) return INFINITY;
    if (z < -1021.0) return 0.0;
    double n = nearbyint(z);
    double g = (z - n) * 0.6931471805599453;
    double q = 1.0 / 6227020800.0;
    q = fma(q, g, 1.0 / 479001600.0);
    q = fma(q, g, 1.0 / 39916800.0);
    q = fma(q, g, 1.0 / 3628800.0);
    q = fma(q, g, 1.0 / 362880.0);
    q = fma(q, g, 1.0 / 40320.0);
    q = fma(q, g, 1.0 / 5040.0);
    q = fma(q, g, 1.0 / 720.0);
    q = fma(q, g, 1.0 / 120.0);
    q = fma(q, g, 1.0 / 24.0);
    q = fma(q, g, 1.0 / 6.0);
    q = fma(q, g, 0.5);
    q = fma(q, g, 1.0);
    q = fma(q, g, 1.0);
    union { double d; uint64_t u; } s;
    s.u = (uint64_t)((int64_t)n + 1023) << 52;
    return q * s.d;
}
/* x^y for x >= 0, y > 0 (what minkowski needs): 0^y = 0, else 2^(y log2 x) */
static double det_pow(double x, double y) { return x > 0.0 ? det_exp2(y * det_log2(x)) : 0.0; }
ORC_API double orc_det_pow(double x, double y) { return det_pow(x, y); }
/* parameter of the metrics that take one (distances.py:138-153 minkowski: p); the oracle is test infrastructure, a process-wide
 * setting is enough */
static double g_metric_param = 2.0;
ORC_API void orc_set_metric_param(double v) { g_metric_param = v; }

/* distances.py:63-91 squared_euclidean (canonical order: sequential fma chain) */
static inline float d_sqeuclid(const float *x, const float *y, int d) {
    float acc = 0.0f;
    for (int e = 0; e < d; ++e) {
        float diff = x[e] - y[e];
        acc = __builtin_fmaf(diff, diff, acc);
    }
    return acc;
}
static inline float seq_dot(const float *x, const float *y, int d) {
    float acc = 0.0f;
    for (int e = 0; e < d; ++e) acc = __builtin_fmaf(x[e], y[e], acc);
    return acc;
}
/* distances.py:583-630 alternative_cosine */
static inline float d_alt_cosine(const float *x, const float *y, int d) {
    float dot = seq_dot(x, y, d), nx = seq_dot(x, x, d), ny = seq_dot(y, y, d);
    if (nx == 0.0f && ny == 0.0f) return 0.0f;
    if (nx == 0.0f || ny == 0.0f) return FLT_MAX;
    if (dot <= 0.0f) return FLT_MAX;
    float r = sqrtf(nx * ny) / dot;
    return (float)det_log2((double)r);
}
/* distances.py:665-701 alternative_dot */
static inline float d_alt_dot(const float *x, const float *y, int d) {
    float dot = seq_dot(x, y, d);
    if (dot <= 0.0f) return FLT_MAX;
    return (float)(-det_log2((double)dot));
}
/* distances.py:759-790 alternative_inner_product: `result` is a float32 local; 1.0 / result is formed in f64 and returned as f4 */
static inline float d_alt_inner(const float *x, const float *y, int d) {
    float dot = seq_dot(x, y, d);
    if (dot <= 0.0f) return FLT_MAX;
    return (float)(1.0 / (double)dot);
}
/* distances.py:138-153 minkowski: result (float64) += (f32 |x_i - y_i|) ** p, then result ** (1 / p) */
static inline float d_minkowski(const float *x, const float *y, int d) {
    const double p = g_metric_param;
    double acc = 0.0;
    for (int e = 0; e < d; ++e) acc += det_pow((double)fabsf(x[e] - y[e]), p);
    return (float)det_pow(acc, 1.0 / p);
}
/* distances.py:1283-1320 correlation (f64 accumulators as in the reference) */
static inline double seq_mean(const float *x, int d) {
    double s = 0.0;
    for (int e = 0; e < d; ++e) s += (double)x[e];
    return s / (double)d;
}
static inline float d_correlation(const float *x, const float *y, int d) {
    double mx = seq_mean(x, d), my = seq_mean(y, d);
    double nx = 0.0, ny = 0.0, dp = 0.0;
    for (int e = 0; e < d; ++e) {
        double sx = (double)x[e] - mx, sy = (double)y[e] - my;
        nx = fma(sx, sx, nx);
        ny = fma(sy, sy, ny);
        dp = fma(sx, sy, dp);
    }
    if (nx == 0.0 && ny == 0.0) return 0.0f;
    if (dp == 0.0) return 1.0f;
    return (float)(1.0 - dp / sqrt(nx * ny));
}
/* distances.py manhattan: the reference's accumulator is a float64 `result`, terms are f32 |x_i - y_i| */
static inline float d_manhattan(const float *x, const float *y, int d) {
    double acc = 0.0;
    for (int e = 0; e < d; ++e) acc += (double)fabsf(x[e] - y[e]);
    return (float)acc;
}
/* distances.py chebyshev */
static inline float d_chebyshev(const float *x, const float *y, int d) {
    float acc = 0.0f;
    for (int e = 0; e < d; ++e) { float t = fabsf(x[e] - y[e]); acc = t > acc ? t : acc; }
    return acc;
}
/* distances.py canberra: f32 quotient |x-y| / (|x|+|y|) added to a float64 result when the denominator is > 0 */
static inline float d_canberra(const float *x, const float *y, int d) {
    double acc = 0.0;
    for (int e = 0; e < d; ++e) {
        float den = fabsf(x[e]) + fabsf(y[e]);
        if (den > 0.0f) acc += (double)(fabsf(x[e] - y[e]) / den);
    }
    return (float)acc;
}
/* distances.py hamming */
static inline float d_hamming(const float *x, const float *y, int d) {
    double acc = 0.0;
    for (int e = 0; e < d; ++e) if (x[e] != y[e]) acc += 1.0;
    return (float)(acc / (double)d);
}
/* the binary family (distances.py jaccard..yule): every one is a function of tt = #{both non-zero}, the two
 * non-zero counts and the dimension; the reference forms the final quotient in float64                        */
static inline float d_binary(int metric, const float *x, const float *y, int d) {
    double tt = 0.0, nx = 0.0, ny = 0.0, n = (double)d;
    for (int e = 0; e < d; ++e) {
        int xt = x[e] != 0.0f, yt = y[e] != 0.0f;
        tt += (double)(xt && yt); nx += (double)xt; ny += (double)yt;
    }
    double neq = nx + ny - 2.0 * tt; /* num_not_equal */
    switch (metric) {
    case ORC_ALT_JACCARD: { /* distances.py:303-326 */
        double nnz = nx + ny - tt;
        if (nnz == 0.0) return 0.0f;
        if (tt == 0.0) return INFINITY; /* -log2(0) */
        return (float)(-det_log2((double)((float)tt / (float)nnz))); /* f32 quotient: numba locals are float32 */
    }
    case ORC_DICE: return neq == 0.0 ? 0.0f : (float)(neq / (2.0 * tt + neq));
    case ORC_MATCHING: return (float)(neq / n);
    case ORC_KULSINSKI: return neq == 0.0 ? 0.0f : (float)((neq - tt + n) / (neq + n));
    case ORC_ROGERSTANIMOTO: return (float)((2.0 * neq) / (n + neq));
    case ORC_RUSSELLRAO: return (tt == nx && tt == ny) ? 0.0f : (float)((n - tt) / n);
    case ORC_SOKALSNEATH: return neq == 0.0 ? 0.0f : (float)(neq / (0.5 * tt + neq));
    default: { /* ORC_YULE */
        double tf = nx - tt, ft = ny - tt, ff = n - tt - tf - ft;
        if (tf == 0.0 || ft == 0.0) return 0.0f;
        return (float)((2.0 * tf * ft) / (tt * ff + tf * ft));
    }
    }
}
/* distances.py:1387-1416 alternative_hellinger: result and both l1 norms are float32 locals in the reference
 * (numba `locals`), summed here in the canonical sequential order; the tail is alternative_cosine's        */
static inline float d_alt_hellinger(const float *x, const float *y, int d) {
    float s = 0.0f, lx = 0.0f, ly = 0.0f;
    for (int e = 0; e < d; ++e) {
        s = s + sqrtf(x[e] * y[e]);
        lx = lx + x[e];
        ly = ly + y[e];
    }
    if (lx == 0.0f && ly == 0.0f) return 0.0f;
    if (lx == 0.0f || ly == 0.0f) return FLT_MAX;
    if (s <= 0.0f) return FLT_MAX;
    float r = sqrtf(lx * ly) / s;
    return (float)det_log2((double)r);
}
static inline float dist_fn(int metric, const float *x, const float *y, int d) {
    switch (metric) {
    case ORC_SQEUCLIDEAN: return d_sqeuclid(x, y, d);
    case ORC_ALT_COSINE: return d_alt_cosine(x, y, d);
    case ORC_ALT_DOT: return d_alt_dot(x, y, d);
    case ORC_CORRELATION: return d_correlation(x, y, d);
    case ORC_MANHATTAN: return d_manhattan(x, y, d);
    case ORC_CHEBYSHEV: return d_chebyshev(x, y, d);
    case ORC_CANBERRA: return d_canberra(x, y, d);
    case ORC_HAMMING: return d_hamming(x, y, d);
    case ORC_ALT_HELLINGER: return d_alt_hellinger(x, y, d);
    case ORC_ALT_INNER: return d_alt_inner(x, y, d);
    case ORC_MINKOWSKI: return d_minkowski(x, y, d);
    default: return d_binary(metric, x, y, d);
    }
}
ORC_API float orc_dist(int metric, const float *x, const float *y, int d) { return dist_fn(metric, x, y, d); }

/* ---------------------------------------------------------------- heaps ---- */
/* utils.py:471-533 checked_flagged_heap_push (flags==NULL -> utils.py:409-468
 * checked_heap_push; check==0 -> utils.py:352-406 simple_heap_push)            */
static inline int heap_push(float *pri, int32_t *ind, uint8_t *flg, int size, float p, int32_t n, uint8_t f,
                            int check) {
    if (p >= pri[0]) return 0;
    if (check)
        for (int i = 0; i < size; ++i)
            if (n == ind[i]) return 0;
    pri[0] = p; ind[0] = n;
    if (flg) flg[0] = f;
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
    return 1;
}
ORC_API int orc_checked_flagged_heap_push(float *pri, int32_t *ind, uint8_t *flg, int size, float p, int32_t n,
                                          uint8_t f) { return heap_push(pri, ind, flg, size, p, n, f, 1); }
ORC_API int orc_checked_heap_push(float *pri, int32_t *ind, int size, float p, int32_t n) {
    return heap_push(pri, ind, NULL, size, p, n, 0, 1);
}
ORC_API int orc_simple_heap_push(float *pri, int32_t *ind, int size, float p, int32_t n) {
    return heap_push(pri, ind, NULL, size, p, n, 0, 0);
}

/* utils.py:165-186 siftdown ; utils.py:189-218 deheap_sort */
static void siftdown(float *h1, int32_t *h2, int size, int elt) {
    while (elt * 2 + 1 < size) {
        int l = elt * 2 + 1, r = l + 1, swap = elt;
        if (h1[swap] < h1[l]) swap = l;
        if (r < size && h1[swap] < h1[r]) swap = r;
        if (swap == elt) break;
        float tf = h1[elt]; h1[elt] = h1[swap]; h1[swap] = tf;
        int32_t ti = h2[elt]; h2[elt] = h2[swap]; h2[swap] = ti;
        elt = swap;
    }
}
ORC_API void orc_deheap_sort(int32_t *ind, float *dist, int64_t n, int k) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        int32_t *ri = ind + i * k;
        float *rd = dist + i * k;
        for (int j = k - 1; j > 0; --j) {
            int32_t ti = ri[0]; ri[0] = ri[j]; ri[j] = ti;
            float tf = rd[0]; rd[0] = rd[j]; rd[j] = tf;
            siftdown(rd, ri, j, 0);
        }
    }
}

/* ------------------------------------------------------- candidate build ---- */
/* utils.py:221-320 new_build_candidates.  T emulated numba threads: thread n owns
 * vertex range [n*B,(n+1)*B), B=N//T+1, stream = rng_state+n (copy; rng_state is
 * never advanced), scans ALL edges in heap-layout order, one draw per touching edge. */
ORC_API void orc_new_build_candidates(const int32_t *ind, uint8_t *flags, int64_t N, int k, int C,
                                      const int64_t *rng_state, int T, int32_t *new_ind, int32_t *old_ind) {
    float *new_pri = (float *)malloc(sizeof(float) * N * C);
    float *old_pri = (float *)malloc(sizeof(float) * N * C);
    for (int64_t i = 0; i < N * C; ++i) { new_ind[i] = -1; old_ind[i] = -1; new_pri[i] = INFINITY; old_pri[i] = INFINITY; }
    int64_t B = N / T + 1;
#pragma omp parallel for schedule(static, 1)
    for (int n = 0; n < T; ++n) {
        int64_t st[3] = {rng_state[0] + n, rng_state[1] + n, rng_state[2] + n};
        int64_t bs = n * B, be = bs + B < N ? bs + B : N;
        for (int64_t i = 0; i < N; ++i) {
            int i_in = (i >= bs && i < be);
            for (int j = 0; j < k; ++j) {
                int64_t idx = ind[i * k + j];
                if (idx < 0) continue;
                int x_in = (idx >= bs && idx < be);
                if (!(i_in || x_in)) continue;
                uint8_t isn = flags[i * k + j];
                float d = orc_tau_rand(st);
                float *P = isn ? new_pri : old_pri;
                int32_t *I = isn ? new_ind : old_ind;
                if (i_in) heap_push(P + i * C, I + i * C, NULL, C, d, (int32_t)idx, 0, 1);
                if (x_in) heap_push(P + idx * C, I + idx * C, NULL, C, d, (int32_t)i, 0, 1);
            }
        }
    }
    /* utils.py:311-318 clear the flag of every entry that made it into new_cand[i] */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i)
        for (int j = 0; j < k; ++j) {
            int32_t idx = ind[i * k + j];
            for (int c = 0; c < C; ++c)
                if (new_ind[i * C + c] == idx) { flags[i * k + j] = 0; break; }
        }
    free(new_pri); free(old_pri);
}

/* ------------------------------------------------- local join + apply ---- */
typedef struct { int32_t p, q; float d; } upd_t;

/* utils.py:536-658 generate_graph_update_array over candidate rows [r0,r1).
 * Emission order = ascending (row, j, new k>=j, old k) (thread partition only
 * chunks that order).  Returns number of updates appended to *out (grown).     */
static int64_t gen_updates_rows(const float *X, int d, int metric, const int32_t *newc, const int32_t *oldc, int C,
                                const float *thr, int64_t r0, int64_t r1, upd_t **out, int64_t *cap) {
    int64_t n = 0;
    for (int64_t i = r0; i < r1; ++i) {
        for (int j = 0; j < C; ++j) {
            int32_t p = newc[i * C + j];
            if (p < 0) continue;
            const float *xp = X + (int64_t)p * d;
            float tp = thr[p];
            for (int kk = j; kk < C; ++kk) {
                int32_t q = newc[i * C + kk];
                if (q < 0) continue;
                float dd = dist_fn(metric, xp, X + (int64_t)q * d, d);
                float mt = tp > thr[q] ? tp : thr[q];
                if (dd <= mt) {
                    if (n == *cap) { *cap = *cap * 2 + 1024; *out = (upd_t *)realloc(*out, sizeof(upd_t) * *cap); }
                    (*out)[n].p = p; (*out)[n].q = q; (*out)[n].d = dd; ++n;
                }
            }
            for (int kk = 0; kk < C; ++kk) {
                int32_t q = oldc[i * C + kk];
                if (q < 0) continue;
                float dd = dist_fn(metric, xp, X + (int64_t)q * d, d);
                float mt = tp > thr[q] ? tp : thr[q];
                if (dd <= mt) {
                    if (n == *cap) { *cap = *cap * 2 + 1024; *out = (upd_t *)realloc(*out, sizeof(upd_t) * *cap); }
                    (*out)[n].p = p; (*out)[n].q = q; (*out)[n].d = dd; ++n;
                }
            }
        }
    }
    return n;
}

/* utils.py:661-733 apply_graph_update_array: owner thread n applies, in list order,
 * (d,q)->row p then (d,p)->row q for rows it owns.  `lists` = T generator buffers
 * scanned in order.                                                             */
static int64_t apply_updates(int32_t *ind, float *dist, uint8_t *flg, int64_t N, int k, upd_t **lists,
                             const int64_t *counts, int nlists, int T) {
    int64_t B = N / T + 1, total = 0;
#pragma omp parallel for schedule(static, 1) reduction(+ : total)
    for (int n = 0; n < T; ++n) {
        int64_t bs = n * B, be = bs + B < N ? bs + B : N;
        for (int t = 0; t < nlists; ++t)
            for (int64_t j = 0; j < counts[t]; ++j) {
                int32_t p = lists[t][j].p, q = lists[t][j].q;
                float dd = lists[t][j].d;
                if (p >= bs && p < be)
                    total += heap_push(dist + (int64_t)p * k, ind + (int64_t)p * k, flg + (int64_t)p * k, k, dd, q, 1, 1);
                if (q >= bs && q < be)
                    total += heap_push(dist + (int64_t)q * k, ind + (int64_t)q * k, flg + (int64_t)q * k, k, dd, p, 1, 1);
            }
    }
    return total;
}

/* pynndescent_.py:217-263 process_candidates: blocks of `block_size` candidate rows,
 * thresholds = heap.dist[:,0] re-read per block.  If upd_out != NULL the emitted
 * updates of ALL blocks are also returned concatenated (test hook).               */
ORC_API int64_t orc_process_candidates(const float *X, int64_t N, int d, int metric, int32_t *ind, float *dist,
                                       uint8_t *flg, int k, const int32_t *newc, const int32_t *oldc, int C,
                                       int64_t block_size, int T, int64_t *n_emitted) {
    int64_t c = 0, emitted = 0;
    int64_t n_blocks = N / block_size;
    upd_t **lists = (upd_t **)calloc(T, sizeof(upd_t *));
    int64_t *caps = (int64_t *)calloc(T, sizeof(int64_t));
    int64_t *counts = (int64_t *)calloc(T, sizeof(int64_t));
    float *thr = (float *)malloc(sizeof(float) * N);
    for (int64_t b = 0; b <= n_blocks; ++b) {
        int64_t bs = b * block_size, be = (b + 1) * block_size < N ? (b + 1) * block_size : N;
        if (bs >= be) continue;
        for (int64_t i = 0; i < N; ++i) thr[i] = dist[i * k];
        int64_t bsz = be - bs, rpt = bsz / T + 1;
#pragma omp parallel for schedule(static, 1)
        for (int t = 0; t < T; ++t) {
            int64_t r0 = bs + t * rpt, r1 = r0 + rpt;
            if (r1 > be) r1 = be;
            counts[t] = (r0 < r1) ? gen_updates_rows(X, d, metric, newc, oldc, C, thr, r0, r1, &lists[t], &caps[t]) : 0;
        }
        for (int t = 0; t < T; ++t) emitted += counts[t];
        c += apply_updates(ind, dist, flg, N, k, lists, counts, T, T);
    }
    for (int t = 0; t < T; ++t) free(lists[t]);
    free(lists); free(caps); free(counts); free(thr);
    if (n_emitted) *n_emitted = emitted;
    return c;
}

/* test hook: emitted updates of one whole-graph join against a given threshold vector */
ORC_API int64_t orc_generate_updates(const float *X, int64_t N, int d, int metric, const int32_t *newc,
                                     const int32_t *oldc, int C, const float *thr, int64_t r0, int64_t r1,
                                     int32_t *out_p, int32_t *out_q, float *out_d, int64_t out_cap) {
    upd_t *l = NULL; int64_t cap = 0;
    int64_t n = gen_updates_rows(X, d, metric, newc, oldc, C, thr, r0, r1, &l, &cap);
    for (int64_t i = 0; i < n && i < out_cap; ++i) { out_p[i] = l[i].p; out_q[i] = l[i].q; out_d[i] = l[i].d; }
    free(l);
    (void)N;
    return n;
}

/* ------------------------------------------------------------- leaf init ---- */
/* pynndescent_.py:73-185 generate_leaf_updates + init_rp_tree: blocks of 8*64 = 512
 * leaves (n_threads is pinned to its default 8, :122 / :343), strict '<' filter,
 * thresholds re-read per block, owner-partitioned apply in emission order.        */
ORC_API void orc_init_rp_tree(const float *X, int64_t N, int d, int metric, int32_t *ind, float *dist, uint8_t *flg,
                              int k, const int32_t *leaf_array, int64_t n_leaves, int max_leaf) {
    const int T = 8;
    int64_t block = T * 64, n_blocks = n_leaves / block;
    upd_t **lists = (upd_t **)calloc(T, sizeof(upd_t *));
    int64_t *caps = (int64_t *)calloc(T, sizeof(int64_t));
    int64_t *counts = (int64_t *)calloc(T, sizeof(int64_t));
    float *thr = (float *)malloc(sizeof(float) * N);
    for (int64_t b = 0; b <= n_blocks; ++b) {
        int64_t bs = b * block, be = (b + 1) * block < n_leaves ? (b + 1) * block : n_leaves;
        int64_t nl = be - bs;
        if (nl <= 0) continue;
        for (int64_t i = 0; i < N; ++i) thr[i] = dist[i * k];
        int64_t lpt = nl / T + 1;
#pragma omp parallel for schedule(static, 1)
        for (int t = 0; t < T; ++t) {
            int64_t l0 = t * lpt, l1 = l0 + lpt < nl ? l0 + lpt : nl;
            int64_t n = 0;
            for (int64_t li = l0; li < l1; ++li) {
                const int32_t *leaf = leaf_array + (bs + li) * max_leaf;
                for (int i = 0; i < max_leaf; ++i) {
                    int32_t p = leaf[i];
                    if (p < 0) break;
                    for (int j = i + 1; j < max_leaf; ++j) {
                        int32_t q = leaf[j];
                        if (q < 0) break;
                        float dd = dist_fn(metric, X + (int64_t)p * d, X + (int64_t)q * d, d);
                        float mt = thr[p] > thr[q] ? thr[p] : thr[q];
                        if (dd < mt) {
                            if (n == caps[t]) { caps[t] = caps[t] * 2 + 1024; lists[t] = (upd_t *)realloc(lists[t], sizeof(upd_t) * caps[t]); }
                            lists[t][n].p = p; lists[t][n].q = q; lists[t][n].d = dd; ++n;
                        }
                    }
                }
            }
            counts[t] = n;
        }
        apply_updates(ind, dist, flg, N, k, lists, counts, T, T);
    }
    for (int t = 0; t < T; ++t) free(lists[t]);
    free(lists); free(caps); free(counts); free(thr);
}

/* utils.py:836-860 initalize_heap_from_graph_indices(_and_distances): warm start from a user-supplied graph
 * (pynndescent_.py:1225-1242).  Row i receives, in column order, a flagged checked push of every j >= 0 with
 * d = metric(data[i], data[j]) -- or the caller's init_dist[i, idx] verbatim when given.                      */
ORC_API void orc_init_from_graph(const float *X, int64_t N, int d, int metric, int32_t *ind, float *dist, uint8_t *flg,
                                 int k, const int32_t *graph, int w, const float *init_dist) {
    for (int64_t i = 0; i < N; ++i)
        for (int c = 0; c < w; ++c) {
            int32_t j = graph[i * w + c];
            if (j < 0) continue;
            float dd = init_dist ? init_dist[i * w + c] : dist_fn(metric, X + i * d, X + (int64_t)j * d, d);
            heap_push(dist + i * k, ind + i * k, flg + i * k, k, dd, j, 1, 1);
        }
}

/* pynndescent_.py:207-214 init_from_neighbor_graph (update(), :2512-2515): row p of the OLD graph (n_rows <= rows of
 * the heap) receives, in column order, an UNFLAGGED checked push of every entry -- invalidated entries are (-1, inf)
 * and fall out at the `d >= root` test of the push itself, exactly as in the reference.                           */
ORC_API void orc_init_from_neighbor_graph(int32_t *ind, float *dist, uint8_t *flg, int k, int64_t n_rows,
                                          const int32_t *graph, const float *gdist, int w) {
    for (int64_t p = 0; p < n_rows; ++p)
        for (int c = 0; c < w; ++c)
            heap_push(dist + p * k, ind + p * k, flg + p * k, k, gdist[p * w + c], graph[p * w + c], 0, 1);
}

/* pynndescent_.py:188-203 init_random (serial; advances rng_state) */
ORC_API void orc_init_random(const float *X, int64_t N, int d, int metric, int32_t *ind, float *dist, uint8_t *flg,
                             int k, int64_t *rng_state) {
    for (int64_t i = 0; i < N; ++i) {
        if (ind[i * k] < 0) {
            int filled = 0;
            for (int j = 0; j < k; ++j) filled += ind[i * k + j] >= 0;
            for (int j = 0; j < k - filled; ++j) {
                int64_t idx = floormod(abs_i32(orc_tau_rand_int(rng_state)), N);
                float dd = dist_fn(metric, X + idx * d, X + i * d, d);
                heap_push(dist + i * k, ind + i * k, flg + i * k, k, dd, (int32_t)idx, 1, 1);
            }
        }
    }
}

/* ------------------------------------------------------------- RP forest ---- */
/* canonical margin arithmetic: 32 interleaved partial sums at float4 granularity
 * (element e belongs to lane (e/4)%32), then an xor butterfly 16,8,4,2,1.          */
static float lane_dot(const float *h, const float *x, int d) {
    float acc[32];
    for (int l = 0; l < 32; ++l) acc[l] = 0.0f;
    for (int e = 0; e < d; ++e) { int l = (e >> 2) & 31; acc[l] = __builtin_fmaf(h[e], x[e], acc[l]); }
    for (int o = 16; o >= 1; o >>= 1) {
        float nw[32];
        for (int l = 0; l < 32; ++l) nw[l] = acc[l] + acc[l ^ o];
        memcpy(acc, nw, sizeof(acc));
    }
    return acc[0];
}
/* utils.py:75-90 norm(): f32 products accumulated in f64 (same lane order), sqrt in f64, returned as f4 */
static float lane_norm(const float *x, int d) {
    double acc[32];
    for (int l = 0; l < 32; ++l) acc[l] = 0.0;
    for (int e = 0; e < d; ++e) { int l = (e >> 2) & 31; acc[l] += (double)(x[e] * x[e]); }
    for (int o = 16; o >= 1; o >>= 1) {
        double nw[32];
        for (int l = 0; l < 32; ++l) nw[l] = acc[l] + acc[l ^ o];
        memcpy(acc, nw, sizeof(acc));
    }
    return (float)sqrt(acc[0]);
}

/* hyperplane through the two chosen points.
 * euclidean: rp_trees.py:358-367 ; angular: rp_trees.py:95-118 */
static void make_hyperplane(const float *X, int d, int angular, int32_t left, int32_t right, float *h, float *off) {
    const float *xl = X + (int64_t)left * d, *xr = X + (int64_t)right * d;
    if (!angular) {
        float o = 0.0f;
        for (int e = 0; e < d; ++e) {
            h[e] = xl[e] - xr[e];
            float t = h[e] * (xl[e] + xr[e]);
            o = o - 0.5f * t;
        }
        *off = o;
    } else {
        float ln = lane_norm(xl, d), rn = lane_norm(xr, d);
        if (fabs((double)ln) < 1e-8) ln = 1.0f;
        if (fabs((double)rn) < 1e-8) rn = 1.0f;
        for (int e = 0; e < d; ++e) h[e] = (xl[e] / ln) - (xr[e] / rn);
        float hn = lane_norm(h, d);
        if (fabs((double)hn) < 1e-8) hn = 1.0f;
        for (int e = 0; e < d; ++e) h[e] = h[e] / hn;
        *off = 0.0f;
    }
}

typedef struct {
    /* flat post-order node store (rp_trees.py:2224-2232 append order) */
    int64_t n_nodes, cap_nodes;
    int32_t *child_l, *child_r; /* -1/-1 for leaves */
    float *offsets;             /* -inf for leaves */
    float *hyper;               /* n_nodes * d (leaf rows unused) */
    int64_t *leaf_start;        /* into leaf_points, -1 for internal */
    int32_t *leaf_size;
    int32_t *leaf_points; int64_t n_leaf_points, cap_leaf_points;
    int d;
} tree_t;

static int64_t tree_add_node(tree_t *t) {
    if (t->n_nodes == t->cap_nodes) {
        t->cap_nodes = t->cap_nodes * 2 + 64;
        t->child_l = (int32_t *)realloc(t->child_l, sizeof(int32_t) * t->cap_nodes);
        t->child_r = (int32_t *)realloc(t->child_r, sizeof(int32_t) * t->cap_nodes);
        t->offsets = (float *)realloc(t->offsets, sizeof(float) * t->cap_nodes);
        t->hyper = (float *)realloc(t->hyper, sizeof(float) * t->cap_nodes * t->d);
        t->leaf_start = (int64_t *)realloc(t->leaf_start, sizeof(int64_t) * t->cap_nodes);
        t->leaf_size = (int32_t *)realloc(t->leaf_size, sizeof(int32_t) * t->cap_nodes);
    }
    return t->n_nodes++;
}

/* rp_trees.py:323-420 euclidean_random_projection_split / :60-171 angular_...  +
 * rp_trees.py:2177-2302 make_euclidean_tree / make_angular_tree (DFS, left first,
 * ONE shared rng stream per tree).                                                */
static void dfs_tree(tree_t *t, const float *X, int d, int angular, int32_t *indices, int64_t n, int64_t *rng,
                     int leaf_size, int max_depth) {
    if (n > leaf_size && max_depth > 0) {
        int64_t li = floormod(orc_tau_rand_int(rng), n);
        int64_t ri = floormod(orc_tau_rand_int(rng), n);
        ri += (li == ri);
        ri = ri % n;
        float *h = (float *)malloc(sizeof(float) * d);
        float off;
        make_hyperplane(X, d, angular, indices[li], indices[ri], h, &off);
        /* The reference accumulates offset and margin in float32 under fastmath=True (rp_trees.py:303-321
         * `locals`), i.e. in whatever order LLVM vectorises them -- unknowable here.  This restatement
         * accumulates the f32 products in float64, which yields the EXACT sign of the margin; every f32
         * summation order reproduces that sign except for points within rounding of the hyperplane, so this is
         * the best predictor of the reference's split (all continuous-data goldens in tests/golden agree; a
         * plain f32 chain misplaced one point of latent_manhattan_1500x20).  On integer / binary data with
         * angular trees margins are EXACTLY zero for many points and the reference's coin flips then depend on
         * its rounding noise: those goldens pin the path from the reference's leaf_array onwards instead.
         * (The "level" scheme below -- the GPU forest's formulation -- keeps its own canonical f32 lane order.) */
        double off64 = 0.0;
        if (!angular) {
            const float *xl = X + (int64_t)indices[li] * d, *xr = X + (int64_t)indices[ri] * d;
            for (int e = 0; e < d; ++e) off64 -= (double)(h[e] * (xl[e] + xr[e])) / 2.0;
            off = (float)off64;
        }
        int8_t *side = (int8_t *)malloc(n);
        int64_t nl = 0, nr = 0;
        for (int64_t i = 0; i < n; ++i) {
            const float *xi = X + (int64_t)indices[i] * d;
            double margin = off64;
            for (int e = 0; e < d; ++e) margin += (double)(h[e] * xi[e]);
            if (fabs(margin) < 1e-8) {
                int32_t r = orc_tau_rand_int(rng);
                side[i] = angular ? (int8_t)floormod(r, 2) : (int8_t)(abs_i32(r) % 2 != 0);
                if (side[i] == 0) nl++; else nr++;
            } else if (margin > 0) { side[i] = 0; nl++; }
            else { side[i] = 1; nr++; }
        }
        if (nl == 0 || nr == 0) {
            nl = nr = 0;
            for (int64_t i = 0; i < n; ++i) {
                side[i] = (int8_t)floormod(orc_tau_rand_int(rng), 2);
                if (side[i] == 0) nl++; else nr++;
            }
        }
        int32_t *L = (int32_t *)malloc(sizeof(int32_t) * (nl ? nl : 1));
        int32_t *R = (int32_t *)malloc(sizeof(int32_t) * (nr ? nr : 1));
        nl = nr = 0;
        for (int64_t i = 0; i < n; ++i) { if (side[i] == 0) L[nl++] = indices[i]; else R[nr++] = indices[i]; }
        free(side);
        dfs_tree(t, X, d, angular, L, nl, rng, leaf_size, max_depth - 1);
        int32_t left_node = (int32_t)(t->n_nodes - 1);
        dfs_tree(t, X, d, angular, R, nr, rng, leaf_size, max_depth - 1);
        int32_t right_node = (int32_t)(t->n_nodes - 1);
        free(L); free(R);
        int64_t id = tree_add_node(t);
        t->child_l[id] = left_node; t->child_r[id] = right_node;
        t->offsets[id] = off;
        memcpy(t->hyper + id * d, h, sizeof(float) * d);
        t->leaf_start[id] = -1; t->leaf_size[id] = 0;
        free(h);
    } else {
        int64_t id = tree_add_node(t);
        t->child_l[id] = -1; t->child_r[id] = -1;
        t->offsets[id] = -INFINITY;
        t->leaf_start[id] = t->n_leaf_points; t->leaf_size[id] = (int32_t)n;
        if (t->n_leaf_points + n > t->cap_leaf_points) {
            t->cap_leaf_points = (t->n_leaf_points + n) * 2 + 64;
            t->leaf_points = (int32_t *)realloc(t->leaf_points, sizeof(int32_t) * t->cap_leaf_points);
        }
        memcpy(t->leaf_points + t->n_leaf_points, indices, sizeof(int32_t) * n);
        t->n_leaf_points += n;
    }
}

typedef struct { int n_trees; tree_t *trees; } forest_t;

/* ---- "level" scheme: the SAME split rule and the same stable partition, but the
 * random numbers of a node come from a stateless hash of (tree seed, level, node
 * ordinal) instead of the DFS-ordered stream -- the formulation the GPU forest uses
 * (DESIGN.md section 5).  Restated here so the CUDA forest has a bit-exact checker. */
static inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline uint64_t tree_seed64(const int64_t *s) {
    return mix64(mix64(mix64((uint64_t)s[0]) ^ (uint64_t)s[1]) ^ (uint64_t)s[2]);
}
static inline uint64_t node_key(uint64_t seed, int level, int64_t ordinal) {
    return mix64(seed ^ mix64(((uint64_t)level << 40) ^ (uint64_t)ordinal));
}

typedef struct { int64_t start, size; } seg_t;

/* builds one tree level by level over a permutation array; leaves are the final
 * segments in position order (== the reference's DFS left-to-right leaf order)    */
static void level_tree(tree_t *t, const float *X, int64_t N, int d, int angular, const int64_t *rng, int leaf_size,
                       int max_depth, uint8_t *seg_flag_out) {
    uint64_t seed = tree_seed64(rng);
    int32_t *perm = (int32_t *)malloc(sizeof(int32_t) * N), *perm2 = (int32_t *)malloc(sizeof(int32_t) * N);
    for (int64_t i = 0; i < N; ++i) perm[i] = (int32_t)i;
    uint8_t *segflag = (uint8_t *)calloc(N + 1, 1); /* 1 where a final/active segment starts */
    segflag[0] = 1;
    seg_t *act = (seg_t *)malloc(sizeof(seg_t)), *nxt = NULL;
    int64_t n_act = 0;
    if (N > leaf_size && max_depth > 0) { act[0].start = 0; act[0].size = N; n_act = 1; }
    float *h = (float *)malloc(sizeof(float) * d);
    int8_t *side = (int8_t *)malloc(N > 0 ? N : 1);
    for (int level = 0; n_act > 0; ++level) {
        nxt = (seg_t *)malloc(sizeof(seg_t) * (2 * n_act));
        int64_t n_nxt = 0;
        for (int64_t o = 0; o < n_act; ++o) {
            int64_t s = act[o].start, n = act[o].size;
            uint64_t key = node_key(seed, level, o);
            int64_t li = (int64_t)(mix64(key ^ 0xA) % (uint64_t)n);
            int64_t ri = (int64_t)(mix64(key ^ 0xB) % (uint64_t)n);
            ri += (li == ri);
            ri = ri % n;
            float off;
            make_hyperplane(X, d, angular, perm[s + li], perm[s + ri], h, &off);
            int64_t nl = 0;
            for (int64_t i = 0; i < n; ++i) {
                float margin = off + lane_dot(h, X + (int64_t)perm[s + i] * d, d);
                if (fabs((double)margin) < 1e-8) side[i] = (int8_t)(mix64(key ^ (0x100 + (uint64_t)i)) & 1);
                else side[i] = margin > 0 ? 0 : 1;
                nl += side[i] == 0;
            }
            if (nl == 0 || nl == n) {
                nl = 0;
                for (int64_t i = 0; i < n; ++i) {
                    side[i] = (int8_t)(mix64(key ^ 0x8000000000000000ULL ^ (uint64_t)i) & 1);
                    nl += side[i] == 0;
                }
            }
            int64_t a = 0, b = 0;
            for (int64_t i = 0; i < n; ++i) {
                if (side[i] == 0) perm2[s + a++] = perm[s + i];
                else perm2[s + nl + b++] = perm[s + i];
            }
            memcpy(perm + s, perm2 + s, sizeof(int32_t) * n);
            if (nl > 0 && nl < n) segflag[s + nl] = 1;
            /* children in position order; a child is active iff size>L and depth allows */
            if (nl > leaf_size && max_depth - (level + 1) > 0) { nxt[n_nxt].start = s; nxt[n_nxt].size = nl; n_nxt++; }
            if (n - nl > leaf_size && max_depth - (level + 1) > 0) { nxt[n_nxt].start = s + nl; nxt[n_nxt].size = n - nl; n_nxt++; }
        }
        free(act); act = nxt; n_act = n_nxt;
    }
    free(act);
    /* emit leaves in position order */
    for (int64_t s = 0; s < N;) {
        int64_t e = s + 1;
        while (e < N && !segflag[e]) ++e;
        int64_t id = tree_add_node(t);
        t->child_l[id] = -1; t->child_r[id] = -1; t->offsets[id] = -INFINITY;
        t->leaf_start[id] = t->n_leaf_points; t->leaf_size[id] = (int32_t)(e - s);
        if (t->n_leaf_points + (e - s) > t->cap_leaf_points) {
            t->cap_leaf_points = (t->n_leaf_points + (e - s)) * 2 + 64;
            t->leaf_points = (int32_t *)realloc(t->leaf_points, sizeof(int32_t) * t->cap_leaf_points);
        }
        memcpy(t->leaf_points + t->n_leaf_points, perm + s, sizeof(int32_t) * (e - s));
        t->n_leaf_points += e - s;
        s = e;
    }
    if (seg_flag_out) memcpy(seg_flag_out, segflag, N);
    free(segflag); free(perm); free(perm2); free(h); free(side);
}

/* rp_trees.py:2815-2888 make_forest (dense).  scheme 0 = reference DFS rng order,
 * scheme 1 = level/node-hash rng.  rng_states: int64[n_trees][3] (copied).       */
ORC_API void *orc_make_forest(const float *X, int64_t N, int d, int angular, const int64_t *rng_states, int n_trees,
                              int leaf_size, int max_depth, int scheme) {
    forest_t *f = (forest_t *)calloc(1, sizeof(forest_t));
    f->n_trees = n_trees;
    f->trees = (tree_t *)calloc(n_trees, sizeof(tree_t));
#pragma omp parallel for schedule(dynamic, 1)
    for (int t = 0; t < n_trees; ++t) {
        tree_t *tr = &f->trees[t];
        tr->d = d;
        int64_t rng[3] = {rng_states[t * 3], rng_states[t * 3 + 1], rng_states[t * 3 + 2]};
        if (scheme == 0) {
            int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * N);
            for (int64_t i = 0; i < N; ++i) idx[i] = (int32_t)i;
            dfs_tree(tr, X, d, angular, idx, N, rng, leaf_size, max_depth);
            free(idx);
        } else {
            level_tree(tr, X, N, d, angular, rng, leaf_size, max_depth, NULL);
        }
    }
    return f;
}
ORC_API void orc_forest_free(void *fp) {
    forest_t *f = (forest_t *)fp;
    for (int t = 0; t < f->n_trees; ++t) {
        tree_t *tr = &f->trees[t];
        free(tr->child_l); free(tr->child_r); free(tr->offsets); free(tr->hyper);
        free(tr->leaf_start); free(tr->leaf_size); free(tr->leaf_points);
    }
    free(f->trees); free(f);
}
/* rp_trees.py:2891-2922 rptree_leaf_array: rows = leaves in node-append order, trees
 * concatenated; width = max(leaf_size, largest leaf).  Two-call protocol.          */
ORC_API void orc_forest_leaf_dims(void *fp, int leaf_size, int64_t *n_leaves, int *max_leaf) {
    forest_t *f = (forest_t *)fp;
    int64_t nl = 0; int ml = leaf_size;
    for (int t = 0; t < f->n_trees; ++t)
        for (int64_t i = 0; i < f->trees[t].n_nodes; ++i)
            if (f->trees[t].child_l[i] < 0) { nl++; if (f->trees[t].leaf_size[i] > ml) ml = f->trees[t].leaf_size[i]; }
    *n_leaves = nl; *max_leaf = ml;
}
ORC_API void orc_forest_leaf_array(void *fp, int max_leaf, int32_t *out) {
    forest_t *f = (forest_t *)fp;
    int64_t row = 0;
    for (int t = 0; t < f->n_trees; ++t) {
        tree_t *tr = &f->trees[t];
        for (int64_t i = 0; i < tr->n_nodes; ++i)
            if (tr->child_l[i] < 0) {
                for (int j = 0; j < max_leaf; ++j) out[row * max_leaf + j] = -1;
                memcpy(out + row * max_leaf, tr->leaf_points + tr->leaf_start[i], sizeof(int32_t) * tr->leaf_size[i]);
                row++;
            }
    }
}
ORC_API int64_t orc_forest_tree_nodes(void *fp, int t) { return ((forest_t *)fp)->trees[t].n_nodes; }
/* raw post-order arrays of one tree (for the search-tree conversion / tests) */
ORC_API void orc_forest_tree_arrays(void *fp, int t, int32_t *child_l, int32_t *child_r, float *offsets, float *hyper,
                                    int32_t *leaf_sizes) {
    tree_t *tr = &((forest_t *)fp)->trees[t];
    memcpy(child_l, tr->child_l, sizeof(int32_t) * tr->n_nodes);
    memcpy(child_r, tr->child_r, sizeof(int32_t) * tr->n_nodes);
    memcpy(offsets, tr->offsets, sizeof(float) * tr->n_nodes);
    memcpy(hyper, tr->hyper, sizeof(float) * tr->n_nodes * tr->d);
    memcpy(leaf_sizes, tr->leaf_size, sizeof(int32_t) * tr->n_nodes);
}

/* ------------------------------------------------------------ NN-descent ---- */
/* pynndescent_.py:266-366 nn_descent_internal / nn_descent (heap must be prepared by
 * the caller: make_heap + init_rp_tree + init_random).  Returns iterations run;
 * c_hist[it] = accepted pushes of iteration it.                                   */
ORC_API int orc_nn_descent_internal(const float *X, int64_t N, int d, int metric, int32_t *ind, float *dist,
                                    uint8_t *flg, int k, int C, const int64_t *rng_state, int T, int n_iters,
                                    double delta, int64_t block_size, int64_t *c_hist) {
    int32_t *newc = (int32_t *)malloc(sizeof(int32_t) * N * C), *oldc = (int32_t *)malloc(sizeof(int32_t) * N * C);
    int it;
    for (it = 0; it < n_iters; ++it) {
        orc_new_build_candidates(ind, flg, N, k, C, rng_state, T, newc, oldc);
        int64_t c = orc_process_candidates(X, N, d, metric, ind, dist, flg, k, newc, oldc, C, block_size, T, NULL);
        if (c_hist) c_hist[it] = c;
        if ((double)c <= delta * (double)k * (double)N) { it++; break; }
    }
    free(newc); free(oldc);
    return it;
}

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------- query walk ---- */
/* pynndescent_.py:1770-1883 search_closure (+ :1716-1727 tree descent, rp_trees.py:2662-2677 select_side,
 * utils.py:352-406 simple_heap_push, utils.py:323-349 visited bitset, CPython heapq on (d, idx) tuples).
 * Canonical arithmetic: margins = offset + lane_dot (as in the forest), distances = the sequential chain.      */
typedef struct { float d; int32_t i; } seed_t;
static inline int seed_lt(seed_t a, seed_t b) { return a.d < b.d || (a.d == b.d && a.i < b.i); }
/* heapq.heappush / heappop (Lib/heapq.py _siftdown / _siftup) */
static void hq_push(seed_t *h, int *n, seed_t x) {
    int pos = (*n)++;
    h[pos] = x;
    while (pos > 0) { int parent = (pos - 1) >> 1; if (seed_lt(x, h[parent])) { h[pos] = h[parent]; pos = parent; } else break; }
    h[pos] = x;
}
static seed_t hq_pop(seed_t *h, int *n) {
    seed_t last = h[--(*n)];
    if (*n == 0) return last;
    seed_t ret = h[0];
    int pos = 0, end = *n, child = 1;
    while (child < end) {                       /* _siftup: bubble the smaller child up to the leaf ... */
        int right = child + 1;
        if (right < end && !seed_lt(h[child], h[right])) child = right;
        h[pos] = h[child]; pos = child; child = 2 * pos + 1;
    }
    h[pos] = last;
    while (pos > 0) { int parent = (pos - 1) >> 1; if (seed_lt(last, h[parent])) { h[pos] = h[parent]; pos = parent; } else break; }
    h[pos] = last;                              /* ... then _siftdown it into place */
    return ret;
}

ORC_API void orc_search(const float *data, int64_t N, int d, int metric, const int32_t *indptr, const int32_t *indices,
                        const float *hyperplanes, const float *offsets, const int32_t *children, const int32_t *tree_indices,
                        const float *Q, int64_t nq, int k, double epsilon, float min_distance, int n_neighbors, int normalize,
                        const int64_t *rng_state, int32_t *out_idx, float *out_dist) {
    int64_t rng[3] = {rng_state[0], rng_state[1], rng_state[2]};   /* :1796 one copy, advanced across queries */
    uint8_t *visited = (uint8_t *)malloc(N / 8 + 1);
    seed_t *seeds = (seed_t *)malloc(sizeof(seed_t) * (N + 16));
    float *q = (float *)malloc(sizeof(float) * d);
    for (int64_t qi = 0; qi < nq; ++qi) {
        float *hp = out_dist + qi * k; int32_t *hi = out_idx + qi * k;
        for (int j = 0; j < k; ++j) { hp[j] = INFINITY; hi[j] = -1; }
        memset(visited, 0, N / 8 + 1);
        const float *qq = Q + qi * d;
        if (normalize) {                                             /* :1806-1811 */
            float s = 0.0f;
            for (int e = 0; e < d; ++e) s = s + qq[e] * qq[e];
            float norm = sqrtf(s);
            if (!(norm > 0.0f)) continue;
            for (int e = 0; e < d; ++e) q[e] = qq[e] / norm;
        } else memcpy(q, qq, sizeof(float) * d);
        int ns = 0;
        int32_t node = 0;                                            /* :1716-1727 */
        while (children[2 * node] > 0) {
            float margin = offsets[node] + lane_dot(hyperplanes + (int64_t)node * d, q, d);
            int side;
            if (fabs((double)margin) < 1e-8) side = (int)(abs_i32(orc_tau_rand_int(rng)) % 2 != 0);
            else side = margin > 0 ? 0 : 1;
            node = children[2 * node + side];
        }
        int32_t b0 = -children[2 * node], b1 = -children[2 * node + 1];
        for (int32_t j = b0; j < b1; ++j) {                          /* :1826-1832 */
            int32_t c = tree_indices[j];
            float dd = dist_fn(metric, q, data + (int64_t)c * d, d);
            heap_push(hp, hi, NULL, k, dd, c, 0, 0);
            seed_t s = {dd, c}; hq_push(seeds, &ns, s);
            visited[c >> 3] |= (uint8_t)(1 << (c & 7));
        }
        int n_random = (k < n_neighbors ? k : n_neighbors) - (b1 - b0);   /* :1835-1847 */
        for (int j = 0; j < n_random; ++j) {
            int32_t c = (int32_t)floormod(abs_i32(orc_tau_rand_int(rng)), N);
            int was = visited[c >> 3] & (1 << (c & 7));
            visited[c >> 3] |= (uint8_t)(1 << (c & 7));
            if (!was) {
                float dd = dist_fn(metric, q, data + (int64_t)c * d, d);
                heap_push(hp, hi, NULL, k, dd, c, 0, 0);
                seed_t s = {dd, c}; hq_push(seeds, &ns, s);
            }
        }
        float bound = (float)((double)hp[0] + epsilon * ((double)hp[0] - (double)min_distance));   /* :1850-1852 */
        if (ns == 0) continue;
        seed_t cur = hq_pop(seeds, &ns);
        while (cur.d < bound) {                                      /* :1857-1881 */
            for (int32_t j = indptr[cur.i]; j < indptr[cur.i + 1]; ++j) {
                int32_t c = indices[j];
                int was = visited[c >> 3] & (1 << (c & 7));
                visited[c >> 3] |= (uint8_t)(1 << (c & 7));
                if (!was) {
                    float dd = dist_fn(metric, q, data + (int64_t)c * d, d);
                    if (dd < bound) {
                        heap_push(hp, hi, NULL, k, dd, c, 0, 0);
                        seed_t s = {dd, c}; hq_push(seeds, &ns, s);
                        bound = (float)((double)hp[0] + epsilon * ((double)hp[0] - (double)min_distance));
                    }
                }
            }
            if (ns == 0) break;
            cur = hq_pop(seeds, &ns);
        }
    }
    free(visited); free(seeds); free(q);
}

/* ------------------------------------------------------- search preparation ---- */
/* rp_trees.py:747-803 get_top_k_hub_indices (k=3): insertion by strictly larger degree, first-come on ties */
static int top3_hubs(const int32_t *indices, int64_t n, const int32_t *deg, int32_t *top) {
    int ak = n < 3 ? (int)n : 3;
    int32_t td[3] = {-1, -1, -1};
    for (int64_t i = 0; i < n; ++i) {
        int32_t dg = deg[indices[i]];
        if (dg > td[ak - 1]) {
            int pos = ak - 1;
            while (pos > 0 && dg > td[pos - 1]) --pos;
            for (int j = ak - 1; j > pos; --j) { td[j] = td[j - 1]; top[j] = top[j - 1]; }
            td[pos] = dg; top[pos] = indices[i];
        }
    }
    return ak;
}

/* rp_trees.py:806-932 euclidean_hub_split / :935-1052 angular_hub_split + :1059-1230 recursive builders.
 * Deterministic (the rng is touched only by the degenerate fallback, whose node then becomes a leaf because its
 * balance is 0 < MIN_SPLIT_BALANCE).                                                                          */
static void hub_tree_rec(tree_t *t, const float *X, int d, int angular, const int32_t *deg, int32_t *indices, int64_t n,
                         int64_t *rng, int leaf_size, int max_depth) {
    int is_leaf = !(n > leaf_size && max_depth > 0);
    float *bh = NULL; float boff = 0.0f; int8_t *bside = NULL; int64_t bnl = 0, bnr = 0;
    if (!is_leaf) {
        int32_t top[3]; int nh = top3_hubs(indices, n, deg, top);
        bh = (float *)calloc(d, sizeof(float)); bside = (int8_t *)calloc(n, 1);
        float *h = (float *)malloc(sizeof(float) * d); int8_t *side = (int8_t *)malloc(n);
        float best_balance = 0.0f;
        for (int hi = 0; hi < nh; ++hi)
            for (int hj = hi + 1; hj < nh; ++hj) {
                float off;
                make_hyperplane(X, d, angular, top[hi], top[hj], h, &off);
                int64_t nl = 0, nr = 0;
                for (int64_t i = 0; i < n; ++i) {
                    float margin = off + lane_dot(h, X + (int64_t)indices[i] * d, d);
                    if ((double)margin > 1e-8) { side[i] = 0; nl++; }
                    else if ((double)margin < -1e-8) { side[i] = 1; nr++; }
                    else { side[i] = (int8_t)(i % 2); if (side[i] == 0) nl++; else nr++; }
                }
                if (nl == 0 || nr == 0) continue;
                float balance = (float)(nl < nr ? nl : nr) / (float)n;
                if (balance > best_balance) {
                    best_balance = balance; bnl = nl; bnr = nr; boff = off;
                    memcpy(bh, h, sizeof(float) * d); memcpy(bside, side, n);
                }
            }
        free(h); free(side);
        if (bnl == 0 || bnr == 0) {           /* fallback: random assignment, balance stays 0 -> leaf */
            for (int64_t i = 0; i < n; ++i) (void)orc_tau_rand_int(rng);
            is_leaf = 1;
        } else if (best_balance < 0.1f) is_leaf = 1;
    }
    if (!is_leaf) {
        int32_t *L = (int32_t *)malloc(sizeof(int32_t) * bnl), *R = (int32_t *)malloc(sizeof(int32_t) * bnr);
        int64_t a = 0, b = 0;
        for (int64_t i = 0; i < n; ++i) { if (bside[i] == 0) L[a++] = indices[i]; else R[b++] = indices[i]; }
        hub_tree_rec(t, X, d, angular, deg, L, bnl, rng, leaf_size, max_depth - 1);
        int32_t left_node = (int32_t)(t->n_nodes - 1);
        hub_tree_rec(t, X, d, angular, deg, R, bnr, rng, leaf_size, max_depth - 1);
        int32_t right_node = (int32_t)(t->n_nodes - 1);
        free(L); free(R);
        int64_t id = tree_add_node(t);
        t->child_l[id] = left_node; t->child_r[id] = right_node; t->offsets[id] = boff;
        memcpy(t->hyper + id * d, bh, sizeof(float) * d);
        t->leaf_start[id] = -1; t->leaf_size[id] = 0;
    } else {
        int64_t id = tree_add_node(t);
        t->child_l[id] = -1; t->child_r[id] = -1; t->offsets[id] = -INFINITY;
        t->leaf_start[id] = t->n_leaf_points; t->leaf_size[id] = (int32_t)n;
        if (t->n_leaf_points + n > t->cap_leaf_points) {
            t->cap_leaf_points = (t->n_leaf_points + n) * 2 + 64;
            t->leaf_points = (int32_t *)realloc(t->leaf_points, sizeof(int32_t) * t->cap_leaf_points);
        }
        memcpy(t->leaf_points + t->n_leaf_points, indices, sizeof(int32_t) * n);
        t->n_leaf_points += n;
    }
    free(bh); free(bside);
}

/* rp_trees.py:2926-2961 recursive_convert (pre-order flat tree) */
static void flat_rec(const tree_t *t, int d, float *hyper, float *offsets, int32_t *children, int32_t *indices,
                     int32_t *node_num, int32_t *leaf_start, int32_t tree_node) {
    int32_t me = *node_num;
    if (t->child_l[tree_node] < 0) {
        int32_t leaf_end = *leaf_start + t->leaf_size[tree_node];
        children[2 * me] = -*leaf_start; children[2 * me + 1] = -leaf_end;
        memcpy(indices + *leaf_start, t->leaf_points + t->leaf_start[tree_node], sizeof(int32_t) * t->leaf_size[tree_node]);
        *leaf_start = leaf_end;
    } else {
        memcpy(hyper + (int64_t)me * d, t->hyper + (int64_t)tree_node * d, sizeof(float) * d);
        offsets[me] = t->offsets[tree_node];
        children[2 * me] = me + 1;
        *node_num = me + 1;
        flat_rec(t, d, hyper, offsets, children, indices, node_num, leaf_start, t->child_l[tree_node]);
        children[2 * me + 1] = *node_num + 1;
        *node_num = *node_num + 1;
        flat_rec(t, d, hyper, offsets, children, indices, node_num, leaf_start, t->child_r[tree_node]);
    }
}

/* rp_trees.py:1233-1309 make_hub_tree + :3019-3049 convert_tree_format.  Two-call protocol: returns a handle. */
ORC_API void *orc_hub_tree(const float *X, int64_t N, int d, int angular, const int32_t *ng_idx, int k, int leaf_size,
                           int max_depth, const int64_t *rng_state, int64_t *n_nodes) {
    int32_t *deg = (int32_t *)calloc(N, sizeof(int32_t));      /* rp_trees.py:714-744 compute_global_degrees */
    for (int64_t i = 0; i < N; ++i)
        for (int j = 0; j < k; ++j) { int32_t v = ng_idx[i * k + j]; if (v >= 0 && v < N) deg[v]++; }
    tree_t *t = (tree_t *)calloc(1, sizeof(tree_t)); t->d = d;
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * N);
    for (int64_t i = 0; i < N; ++i) idx[i] = (int32_t)i;
    int64_t rng[3] = {rng_state[0], rng_state[1], rng_state[2]};
    hub_tree_rec(t, X, d, angular, deg, idx, N, rng, leaf_size, max_depth);
    free(idx); free(deg);
    *n_nodes = t->n_nodes;
    return t;
}
ORC_API void orc_hub_tree_flatten(void *tp, int d, int64_t N, float *hyper, float *offsets, int32_t *children, int32_t *indices) {
    tree_t *t = (tree_t *)tp;
    for (int64_t i = 0; i < t->n_nodes; ++i) { offsets[i] = 0.0f; children[2 * i] = -1; children[2 * i + 1] = -1; }
    memset(hyper, 0, sizeof(float) * t->n_nodes * d);
    for (int64_t i = 0; i < N; ++i) indices[i] = -1;
    int32_t node_num = 0, leaf_start = 0;
    flat_rec(t, d, hyper, offsets, children, indices, &node_num, &leaf_start, (int32_t)(t->n_nodes - 1));
    free(t->child_l); free(t->child_r); free(t->offsets); free(t->hyper); free(t->leaf_start); free(t->leaf_size); free(t->leaf_points);
    free(t);
}

/* pynndescent_.py:369-403 diversify (prune_probability = 1: the racy tau_rand draw is always < 1) */
ORC_API void orc_diversify(int32_t *ind, float *dist, int64_t N, int k, const float *X, int d, int metric) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        int32_t *ri = ind + i * k; float *rd = dist + i * k;
        int32_t ni[256]; float nd[256]; int nn = 0;
        ni[nn] = ri[0]; nd[nn] = rd[0]; nn++;
        for (int j = 1; j < k; ++j) {
            if (ri[j] < 0) break;
            int flag = 1;
            for (int c = 0; c < nn; ++c) {
                float dd = dist_fn(metric, X + (int64_t)ri[j] * d, X + (int64_t)ni[c] * d, d);
                if (nd[c] > FLT_EPSILON && dd < rd[j]) { flag = 0; break; }
            }
            if (flag) { ni[nn] = ri[j]; nd[nn] = rd[j]; nn++; }
        }
        for (int j = 0; j < k; ++j) { if (j < nn) { ri[j] = ni[j]; rd[j] = nd[j]; } else { ri[j] = -1; rd[j] = INFINITY; } }
    }
}

/* pynndescent_.py:549-588 diversify_csr, literally (note the reference indexes the second point with k, not order[k]).
 * argsort ties: stable order here (numba's quicksort tie order is unspecified).                                   */
ORC_API void orc_diversify_csr(const int32_t *indptr, const int32_t *indices, float *data, int64_t n_nodes, const float *X,
                               int d, int metric) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n_nodes; ++i) {
        int32_t s = indptr[i], m = indptr[i + 1] - indptr[i];
        if (m <= 1) continue;
        const int32_t *ci = indices + s; float *cd = data + s;
        int32_t *order = (int32_t *)malloc(sizeof(int32_t) * m); int8_t *ret = (int8_t *)malloc(m);
        for (int a = 0; a < m; ++a) { order[a] = a; ret[a] = 1; }
        for (int a = 1; a < m; ++a) { int32_t v = order[a]; int b = a - 1; while (b >= 0 && cd[order[b]] > cd[v]) { order[b + 1] = order[b]; --b; } order[b + 1] = v; }
        for (int a = 1; a < m; ++a) {
            int32_t j = order[a];
            for (int kk = 0; kk < a; ++kk) {
                int32_t l = order[kk];
                if (ret[l] == 1) {
                    float dd = dist_fn(metric, X + (int64_t)ci[j] * d, X + (int64_t)ci[kk] * d, d);
                    if (cd[l] > FLT_EPSILON && dd < cd[j]) { ret[j] = 0; break; }
                }
            }
        }
        for (int a = 0; a < m; ++a) if (ret[order[a]] == 0) cd[order[a]] = 0.0f;
        free(order); free(ret);
    }
}
