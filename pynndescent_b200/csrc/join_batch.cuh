// join_batch.cuh -- parameter block and batching constants of the batched local join (k_join_batch, k_descent.cu).
#pragma once
#include "nnd_common.cuh"

#define TILE 4
#define JB_THREADS 128
#define JB_TCAP 128      // tiles per batch
#define JB_SCAP 176      // physical member slots per batch (2 buffers x 176 x 144 B -> 4 CTAs/SM)
#define JB_MAXU 48       // units per batch (every unit weighs >= 4)
#define JB_CMAX 60  // members scratch is sized for the largest C so that its offset does not depend on C
struct JoinBatchParams {
    const float *X; int64_t N; int ld; int metric; const void *aux; const float *thr;
    const int32_t *act_unit; const uint32_t *act_info; const uint32_t *act_poff; const int32_t *members;
    const int32_t *batch_first;
    int64_t n_active, n_batches; int64_t batch_begin, batch_end; int slot_stride;
    int32_t *out_tgt, *out_oth; float *out_d; uint32_t *out_pos; unsigned long long *out_seq;
    uint32_t *tgt_count; unsigned long long *cursor; unsigned long long capacity; unsigned long long *stats; int *overflow;
};

__host__ __device__ __forceinline__ int jb_tiles(int nn, int m) {
    int ntr = (nn + 3) >> 2, ntc = (m + 3) >> 2;
    return ntr * ntc - ntr * (ntr - 1) / 2;
}
