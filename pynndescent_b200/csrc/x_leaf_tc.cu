// x_leaf_tc.cu -- EXPERIMENT (not on any product path; called only by scripts/leaf_tc_experiment.py): the leaf join's distance
// matrix (pynndescent_.py:73-113 generate_leaf_updates) as a tensor-core Gram with BF16 inputs, used the only way the parity
// contract allows -- as a CONSERVATIVE PREFILTER in front of the exact FP32 difference-form chain (heaps must
// hold the canonical distances bit for bit, so a rounded Gram may only decide which pairs need not be evaluated exactly).
//
//   G = A A^T on mma.sync.m16n8k16 (bf16 x bf16 -> f32), A = the leaf's <= 80 member rows staged in shared memory as bf16;
//   lower bound  lb(i,j) = n_i + n_j - 2 G_ij - 2 * 2^-8 * 1.01 * sqrt(n_i n_j) - 1e-5 (n_i + n_j)  <=  d^2(i,j)
//   (bf16 round-to-nearest: |x^ - x| <= 2^-9 |x| per element => |G - <x,y>| <= (2^-8 + 2^-18) |x||y| by Cauchy-Schwarz; the last
//   term covers the f32 norms / accumulation);  a pair SURVIVES iff lb < max(thr_i, thr_j)  (the leaf filter is d < max(thr), :104).
//
// The kernel counts, per launch: pairs, survivors, and (EXACT = 1) the pairs whose exact distance passes the filter -- the
// survivors must contain all of them (conservativeness is asserted by the script).  What it measures: the tensor-core Gram time
// and the survivor fraction per tree, i.e. how much exact work a prefilter could remove at best.
#include "nnd_common.cuh"
#include <cuda_bf16.h>

#define XL_MP 80          // members per leaf, padded to 5 x 16
#define XL_LD 136         // bf16 row stride (128 + 8: conflict-free 32-bit fragment loads)
#define XL_T 128

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int EXACT>
__global__ void __launch_bounds__(XL_T) k_x_leaf_gram_tc(const float *X, int ld, int d, const int32_t *leaf_points, const int64_t *leaf_offsets,
                                                         int64_t l0, int64_t l1, const float *heap_dist, int k, unsigned long long *counts) {
    __shared__ __align__(16) __nv_bfloat16 As[XL_MP * XL_LD];
    __shared__ float s_n[XL_MP], s_thr[XL_MP];
    __shared__ int32_t s_id[XL_MP];
    __shared__ unsigned int s_cnt[3];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t leaf = l0 + blockIdx.x;
    if (leaf >= l1) return;
    const int64_t o0 = leaf_offsets[leaf];
    const int m = (int)(leaf_offsets[leaf + 1] - o0);
    if (tid < 3) s_cnt[tid] = 0;
    if (tid < XL_MP) {
        const int32_t id = tid < m ? leaf_points[o0 + tid] : -1;
        s_id[tid] = id;
        s_thr[tid] = id >= 0 ? heap_dist[(int64_t)id * k] : 0.0f;
    }
    __syncthreads();
    // rows -> bf16 in shared memory (one warp per row, a float4 per lane), exact f32 squared norms on the way
    for (int r = warp; r < XL_MP; r += XL_T / 32) {
        const int32_t id = s_id[r];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (id >= 0 && lane * 4 < ld) v = *(const float4 *)(X + (int64_t)id * ld + lane * 4);
        __nv_bfloat162 p0 = __floats2bfloat162_rn(v.x, v.y), p1 = __floats2bfloat162_rn(v.z, v.w);
        *(__nv_bfloat162 *)(As + r * XL_LD + lane * 4) = p0;
        *(__nv_bfloat162 *)(As + r * XL_LD + lane * 4 + 2) = p1;
        float n = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        for (int s = 16; s >= 1; s >>= 1) n += __shfl_xor_sync(0xffffffffu, n, s);
        if (lane == 0) s_n[r] = n;
    }
    __syncthreads();
    // upper-triangular 16 x 8 tiles of the 80 x 80 Gram, round-robin over the warps
    const int g = lane >> 2, t = lane & 3;
    unsigned int surv = 0, pairs = 0;
    int tile = 0;
    for (int rt = 0; rt < XL_MP / 16; ++rt)
        for (int ct = 0; ct < XL_MP / 8; ++ct) {
            if (ct * 8 + 7 < rt * 16) continue;             // entirely below the diagonal
            if (rt * 16 >= m || ct * 8 >= m) continue;
            if ((tile++ & 3) != warp) continue;
            float c[4] = {0.f, 0.f, 0.f, 0.f};
            const __nv_bfloat16 *ar = As + (rt * 16 + g) * XL_LD, *br = As + (ct * 8 + g) * XL_LD;
#pragma unroll
            for (int k0 = 0; k0 < 128; k0 += 16) {
                const uint32_t a0 = *(const uint32_t *)(ar + k0 + 2 * t), a1 = *(const uint32_t *)(ar + 8 * XL_LD + k0 + 2 * t);
                const uint32_t a2 = *(const uint32_t *)(ar + k0 + 8 + 2 * t), a3 = *(const uint32_t *)(ar + 8 * XL_LD + k0 + 8 + 2 * t);
                const uint32_t b0 = *(const uint32_t *)(br + k0 + 2 * t), b1 = *(const uint32_t *)(br + k0 + 8 + 2 * t);
                mma_bf16_16816(c, a0, a1, a2, a3, b0, b1);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = rt * 16 + g + (e >> 1) * 8, j = ct * 8 + 2 * t + (e & 1);
                if (i < j && j < m) {
                    const float ni = s_n[i], nj = s_n[j];
                    const float lb = ni + nj - 2.0f * c[e] - 2.0f * 0.00390625f * 1.01f * sqrtf(ni * nj) - 1e-5f * (ni + nj);
                    ++pairs;
                    surv += lb < fmaxf(s_thr[i], s_thr[j]);
                }
            }
        }
    for (int s = 16; s >= 1; s >>= 1) { surv += __shfl_xor_sync(0xffffffffu, surv, s); pairs += __shfl_xor_sync(0xffffffffu, pairs, s); }
    if (lane == 0) { atomicAdd(&s_cnt[0], pairs); atomicAdd(&s_cnt[1], surv); }
    if (EXACT) {
        unsigned int pass = 0;
        for (int e = tid; e < m * m; e += XL_T) {
            const int i = e / m, j = e - i * m;
            if (i >= j) continue;
            const float *xi = X + (int64_t)s_id[i] * ld, *xj = X + (int64_t)s_id[j] * ld;
            float acc = 0.0f;
            for (int q = 0; q < d; ++q) { const float df = __fsub_rn(xi[q], xj[q]); acc = __fmaf_rn(df, df, acc); }   // the canonical chain
            pass += acc < fmaxf(s_thr[i], s_thr[j]);
        }
        for (int s = 16; s >= 1; s >>= 1) pass += __shfl_xor_sync(0xffffffffu, pass, s);
        if (lane == 0) atomicAdd(&s_cnt[2], pass);
    }
    __syncthreads();
    if (tid < 3 && s_cnt[tid]) atomicAdd(&counts[tid], (unsigned long long)s_cnt[tid]);
}

// counts_host[3] = {pairs, survivors of the bf16 bound, pairs whose exact distance passes (exact != 0 only)}; squared-euclidean
// metric, d <= 128, leaves of at most 80 points
extern "C" int nnd_x_leaf_gram_tc(nnd_ctx *ctx, const nnd_problem *P, const int32_t *leaf_points, const int64_t *leaf_offsets, int64_t l0,
                                  int64_t l1, int32_t max_leaf, int32_t exact, unsigned long long *counts_dev, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    NND_CHECK(ctx && P && leaf_points && leaf_offsets && counts_dev && l1 >= l0, NND_ERR_INVALID, "nnd_x_leaf_gram_tc: bad args");
    NND_CHECK(P->metric == NND_METRIC_SQEUCLIDEAN && P->ld <= 128 && max_leaf <= XL_MP, NND_ERR_UNSUPPORTED,
              "nnd_x_leaf_gram_tc: experiment covers squared euclidean, d <= 128, leaves of <= %d points", XL_MP);
    if (l1 == l0) return NND_OK;
    if (exact) k_x_leaf_gram_tc<1><<<(unsigned)(l1 - l0), XL_T, 0, st>>>(P->X, P->ld, P->d, leaf_points, leaf_offsets, l0, l1, P->heap_dist, P->k, counts_dev);
    else k_x_leaf_gram_tc<0><<<(unsigned)(l1 - l0), XL_T, 0, st>>>(P->X, P->ld, P->d, leaf_points, leaf_offsets, l0, l1, P->heap_dist, P->k, counts_dev);
    NND_LAUNCH(ctx);
    NND_CUDA(cudaGetLastError());
    return NND_OK;
}
