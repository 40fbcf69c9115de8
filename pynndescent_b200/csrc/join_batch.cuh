// join_batch.cuh -- definitions shared by the two batched local-join kernels (k_join_batch in k_descent.cu, the
// persistent warp-specialised k_join_pipe in k_join_pipe.cu) and their host launcher.
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

// ---- k_join_pipe (k_join_pipe.cu): per-batch descriptor built by k_jp_meta and pulled into shared memory by ONE bulk copy
#define JP_REC_BYTES 2704
struct __align__(16) JpRec {
    int32_t hdr[4];               // [0] units in the batch (0 = empty weight window) [1] tiles [2] members [3] physical slots
    int32_t ids[JB_SCAP];         // member ids in logical order (unit offset + member position), -1 = pad
    float thr[JB_SCAP];           // heap-root snapshot of every member
    float n2[JB_SCAP];            // cosine: squared norm of every member
    int32_t uid[JB_MAXU];         // candidate row (vertex) of every unit
    uint8_t unn[JB_MAXU], um[JB_MAXU];          // new members / all members per unit
    uint8_t uoff[JB_MAXU + 4], utile[JB_MAXU + 4];   // first physical slot / first tile of every unit (+ end sentinel)
    uint8_t dst[JB_SCAP];         // physical shared-memory slot of logical position e (255 = pad)
};
static_assert(sizeof(JpRec) == JP_REC_BYTES, "JpRec layout");
static_assert(JB_SCAP < 255 && JB_TCAP < 255, "JpRec packs slots and tiles into bytes");

struct JoinPipeOpts { int dk; int stages; int ctas_per_sm; };
int launch_join_pipe(nnd_ctx *ctx, const JoinBatchParams &JB, void *meta_buf, size_t meta_bytes, const JoinPipeOpts &opt,
                     cudaStream_t st);
size_t join_pipe_meta_bytes(int64_t n_batches);
