// k_join_pipe.cu -- persistent, warp-specialised variant of the batched local join (generate_graph_update_array,
// utils.py:536-658).  Same batches, same tiles, same arithmetic order and the same emission keys as k_join_batch
// (k_descent.cu) -- only the schedule differs, so the two kernels are interchangeable bit for bit:
//
//   k_jp_meta   (one warp per batch, ordinary streaming kernel) resolves everything a batch needs that sits behind a
//               chain of dependent global loads -- batch bounds -> unit table -> member ids -> thresholds / norms ->
//               shared-memory slot of every member -- into ONE 2.7 KB descriptor (JpRec, join_batch.cuh).
//   k_join_pipe (persistent: grid = SMs x CTAs/SM, each CTA walks batches b, b+grid, ...): a PRODUCER warp pulls
//               descriptors into a 4-deep ring with one bulk copy each (prefetch distance 2) and gathers the member
//               rows, d-chunk by d-chunk, into an NST-deep ring of stage buffers with the TMA engine's 1-D bulk copy
//               (full/empty mbarrier pairs, SASS UBLKCP + SYNCS); four CONSUMER warps own one 4x4 register tile each,
//               accumulate chunk after chunk in the canonical sequential order and emit the half-updates.  The
//               gathers of batch b+1 are in flight while batch b is being accumulated and there is no CTA-wide
//               barrier in the steady state (k_join_batch: 5 dependent loads before the first gather, one
//               __syncthreads per chunk; profiles/README.md: long_scoreboard 21 %, barrier 20 %).
//
// STATUS: written at the end of round 1 after the GPU budget was spent -- compiled for sm_100a, NOT yet run on
// hardware.  It is therefore OFF by default (NND_JOIN_PIPE=1 turns it on; scripts/r2_ab.py is the A/B harness that
// checks it bit for bit against k_join_batch and times both).
#include "join_batch.cuh"

#define JP_CWARPS 4
#define JP_CONS (JP_CWARPS * 32)
#define JP_THREADS (JP_CONS + 32)
#define JP_NM 4           // descriptor ring
#define JP_MDIST 2        // descriptor prefetch distance (< JP_NM so that the refill never waits for the batch in flight)

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------------------------------------ descriptor prepass
#define JPM_WARPS 4
template <bool COSINE>
__global__ void __launch_bounds__(JPM_WARPS * 32) k_jp_meta(JoinBatchParams P, JpRec *out) {
    __shared__ JpRec srec[JPM_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t B = P.batch_begin + (int64_t)blockIdx.x * JPM_WARPS + warp;
    if (B >= P.batch_end) return;
    JpRec &M = srec[warp];
    JpRec *dstrec = out + (B - P.batch_begin);
    // batch_first[b] = first active row whose weight prefix starts in window b, or a sentinel (>= n_active) when a heavy
    // row of an earlier window covers window b entirely
    const int64_t j0 = P.batch_first[B];
    int64_t j1 = P.n_active;
    if (j0 < P.n_active)
        for (int64_t bb = B + 1; bb < P.n_batches; ++bb) { int64_t v = P.batch_first[bb]; if (v < P.n_active) { j1 = v; break; } }
    const int nunits = j0 < P.n_active ? (int)(j1 - j0) : 0;
    if (nunits <= 0 || nunits > JB_MAXU) {          // empty weight window
        if (lane < 4) dstrec->hdr[lane] = 0;
        if (nunits > JB_MAXU && lane == 0) *P.overflow = 1;   // cannot happen for weight-cut batches; never drop work silently
        return;
    }
    const uint32_t base = P.act_poff[j0];
    int carry_t = 0, memsum = 0;
    unsigned long long pairs = 0;
    for (int u0 = 0; u0 < nunits; u0 += 32) {
        const int u = u0 + lane;
        int nn = 0, m = 0, t = 0;
        if (u < nunits) {
            const uint32_t inf = P.act_info[j0 + u];
            nn = (int)(inf & 0xFF); m = (int)(inf >> 8);
            M.uid[u] = P.act_unit[j0 + u]; M.unn[u] = (uint8_t)nn; M.um[u] = (uint8_t)m;
            M.uoff[u] = (uint8_t)(P.act_poff[j0 + u] - base);
            t = jb_tiles(nn, m);
        }
        int incl = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (u < nunits) M.utile[u] = (uint8_t)(carry_t + incl - t);
        carry_t += __shfl_sync(0xffffffffu, incl, 31);
        int ms = m, pp = nn * (nn + 1) / 2 + nn * (m - nn);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) { ms += __shfl_xor_sync(0xffffffffu, ms, o); pp += __shfl_xor_sync(0xffffffffu, pp, o); }
        memsum += ms; pairs += (unsigned long long)pp;
    }
    __syncwarp();
    const int totphys = (int)M.uoff[nunits - 1] + ((((int)M.um[nunits - 1] + 3) >> 2) << 2);
    if (lane == 0) {
        M.utile[nunits] = (uint8_t)carry_t; M.uoff[nunits] = (uint8_t)totphys;
        M.hdr[0] = nunits; M.hdr[1] = carry_t; M.hdr[2] = memsum; M.hdr[3] = totphys;
        if (P.stats) { atomicAdd(&P.stats[0], (unsigned long long)memsum); atomicAdd(&P.stats[1], pairs); }   // profile counters
    }
    for (int e = lane; e < totphys; e += 32) {
        const int32_t id = P.members[base + e];
        M.ids[e] = id;
        M.thr[e] = id >= 0 ? P.thr[id] : 0.0f;
        if (COSINE) M.n2[e] = id >= 0 ? ((const float *)P.aux)[id] : 0.0f;
    }
    // member j of a unit lives at shared-memory slot (j&3)*ntc + (j>>2) of the unit's slot range (k_join_batch's
    // bank-conflict-free permutation)
    for (int uu = 0; uu < nunits; ++uu) {
        const int mm = M.um[uu], nt = (mm + 3) >> 2, o = M.uoff[uu];
        for (int j = lane; j < nt * 4; j += 32) M.dst[o + j] = j < mm ? (uint8_t)(o + (j & 3) * nt + (j >> 2)) : (uint8_t)255;
    }
    __syncwarp();
    // coalesced copy-out (fields beyond the used prefix are never read by the join)
    const uint4 *s4 = (const uint4 *)&M;
    uint4 *d4 = (uint4 *)dstrec;
    for (int q = lane; q < JP_REC_BYTES / 16; q += 32) d4[q] = s4[q];
}

// ------------------------------------------------------------------------------------------------ the join
template <int DKF, int NST>
__host__ __device__ constexpr size_t jp_smem_bytes() {
    return 256 /* barriers */ + (size_t)JP_NM * JP_REC_BYTES + (size_t)NST * JB_SCAP * (DKF + 4) * 4 + 16;
}

template <int METRIC, int DKF, int NST>
__global__ void __launch_bounds__(JP_THREADS) k_join_pipe(JoinBatchParams P, const JpRec *__restrict__ recs) {
    constexpr int SDF = DKF + 4;                         // consecutive slots shift by 4 banks
    static_assert(2 * NST + 2 * JP_NM <= 32, "barrier block is 256 bytes");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *full = (uint64_t *)smem_raw;               // [NST]    stage filled   (1 arrival + tx bytes)
    uint64_t *empty = full + NST;                         // [NST]    stage released (JP_CWARPS arrivals)
    uint64_t *mfull = empty + NST;                        // [JP_NM]  descriptor landed
    uint64_t *mempty = mfull + JP_NM;                     // [JP_NM]  descriptor released
    JpRec *meta = (JpRec *)(smem_raw + 256);
    float *stages = (float *)(smem_raw + 256 + (size_t)JP_NM * JP_REC_BYTES);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], JP_CWARPS); }
        for (int b = 0; b < JP_NM; ++b) { mbar_init(&mfull[b], 1); mbar_init(&mempty[b], JP_CWARPS); }
        mbar_fence_init();
    }
    __syncthreads();
    const int64_t nb = P.batch_end - P.batch_begin;
    const int nmy = (int64_t)blockIdx.x < nb ? (int)((nb - 1 - blockIdx.x) / gridDim.x) + 1 : 0;   // batches b = blockIdx.x + i*grid
    const int nchunks = (P.ld + DKF - 1) / DKF;

    if (warp == JP_CWARPS) {
        // ================================================================ producer warp
        auto fetch_meta = [&](int i) {      // descriptor of my i-th batch -> ring slot i % JP_NM
            const int q = i % JP_NM, n = i / JP_NM;
            if (n >= 1) mbar_wait(&mempty[q], (uint32_t)((n - 1) & 1));
            if (lane == 0) {
                mbar_expect_tx(&mfull[q], JP_REC_BYTES);
                bulk_g2s(&meta[q], recs + ((int64_t)blockIdx.x + (int64_t)i * gridDim.x), JP_REC_BYTES, &mfull[q]);
            }
        };
        for (int i = 0; i < JP_MDIST && i < nmy; ++i) fetch_meta(i);
        uint32_t it = 0;                                  // stages issued so far
        for (int i = 0; i < nmy; ++i) {
            if (i + JP_MDIST < nmy) fetch_meta(i + JP_MDIST);
            const int q = i % JP_NM;
            mbar_wait(&mfull[q], (uint32_t)((i / JP_NM) & 1));
            const JpRec &M = meta[q];
            const int nunits = M.hdr[0], mem = M.hdr[2], totphys = M.hdr[3];
            // every lane has read the header before lane 0 may run ahead: the consumers release an EMPTY batch without
            // waiting for anything this warp does, so its ring slot could otherwise be refilled (two iterations from
            // now) under a lane that has not looked yet (found by the CPU emulation, tests/test_kernel_emulation.py)
            __syncwarp();
            if (nunits <= 0) continue;
            for (int c = 0; c < nchunks; ++c, ++it) {
                const int s = (int)(it % NST); const uint32_t n = it / NST;
                const int c0 = c * DKF, len = min(DKF, P.ld - c0);
                if (n >= 1) mbar_wait(&empty[s], (n - 1) & 1);
                if (lane == 0) mbar_expect_tx(&full[s], (uint32_t)mem * (uint32_t)len * 4u);
                __syncwarp();
                float *buf = stages + (size_t)s * JB_SCAP * SDF;
                for (int e = lane; e < totphys; e += 32) {
                    const int d = M.dst[e];
                    if (d != 255)
                        bulk_g2s(buf + (size_t)d * SDF, P.X + (int64_t)M.ids[e] * P.ld + c0, (uint32_t)len * 4u, &full[s]);
                }
            }
        }
        return;
    }

    // ==================================================================== consumer warps
    uint32_t it = 0;
    for (int i = 0; i < nmy; ++i) {
        const int q = i % JP_NM;
        mbar_wait(&mfull[q], (uint32_t)((i / JP_NM) & 1));
        const JpRec &M = meta[q];
        const int nunits = M.hdr[0], T = M.hdr[1];
        if (nunits > 0) {
            // ---- my tile: (ta,tb), tb >= ta, of unit u
            int u = -1, ta = 0, tb = 0, nn = 0, m = 0, uo = 0, ntc = 1;
            if (tid < T) {
                int lo = 0, hi = nunits;                    // last u with utile[u] <= tid
                while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if ((int)M.utile[mid] <= tid) lo = mid; else hi = mid; }
                u = lo;
                nn = M.unn[u]; m = M.um[u]; uo = M.uoff[u]; ntc = (m + 3) >> 2;
                int lt = tid - (int)M.utile[u];
                while (lt >= ntc - ta) { lt -= ntc - ta; ++ta; }
                tb = ta + lt;
            }
            int oa[TILE], ob[TILE];
#pragma unroll
            for (int r = 0; r < TILE; ++r) {
                const int ia = min(ta * TILE + r, max(nn - 1, 0)), ib = min(tb * TILE + r, max(m - 1, 0));
                oa[r] = (uo + (ia & 3) * ntc + (ia >> 2)) * SDF;
                ob[r] = (uo + (ib & 3) * ntc + (ib >> 2)) * SDF;
            }
            float acc[TILE * TILE];
#pragma unroll
            for (int p = 0; p < TILE * TILE; ++p) acc[p] = 0.0f;
            for (int c = 0; c < nchunks; ++c, ++it) {
                const int s = (int)(it % NST);
                mbar_wait(&full[s], (it / NST) & 1);
                if (u >= 0) {
                    const float *buf = stages + (size_t)s * JB_SCAP * SDF;
                    const int len = min(DKF, P.ld - c * DKF);
#pragma unroll 2
                    for (int e = 0; e < len; e += 4) {
                        float4 va[TILE], vb[TILE];
#pragma unroll
                        for (int r = 0; r < TILE; ++r) { va[r] = *(const float4 *)(buf + oa[r] + e); vb[r] = *(const float4 *)(buf + ob[r] + e); }
#pragma unroll
                        for (int w = 0; w < 4; ++w) {
#pragma unroll
                            for (int qa = 0; qa < TILE; ++qa) {
                                const float xa = (w == 0) ? va[qa].x : (w == 1) ? va[qa].y : (w == 2) ? va[qa].z : va[qa].w;
#pragma unroll
                                for (int qb = 0; qb < TILE; ++qb) {
                                    const float xb = (w == 0) ? vb[qb].x : (w == 1) ? vb[qb].y : (w == 2) ? vb[qb].z : vb[qb].w;
                                    if (METRIC == NND_METRIC_SQEUCLIDEAN) {
                                        const float df = __fsub_rn(xa, xb);
                                        acc[qa * TILE + qb] = __fmaf_rn(df, df, acc[qa * TILE + qb]);
                                    } else {
                                        acc[qa * TILE + qb] = __fmaf_rn(xa, xb, acc[qa * TILE + qb]);
                                    }
                                }
                            }
                        }
                    }
                }
                __syncwarp();                               // every lane is done reading the stage
                if (lane == 0) mbar_arrive(&empty[s]);
            }
            // ---- finalise + filter (identical to k_join_batch)
            uint32_t pm = 0;
            if (u >= 0) {
#pragma unroll
                for (int qa = 0; qa < TILE; ++qa)
#pragma unroll
                    for (int qb = 0; qb < TILE; ++qb) {
                        const int a = ta * TILE + qa, b = tb * TILE + qb;
                        const bool valid = (a < nn) && (b < m) && (b >= a);
                        float r = 0.0f;
                        if (valid) {
                            if (METRIC == NND_METRIC_SQEUCLIDEAN) r = acc[qa * TILE + qb];
                            else if (METRIC == NND_METRIC_ALT_COSINE) r = finish_alt_cosine(acc[qa * TILE + qb], M.n2[uo + a], M.n2[uo + b]);
                            else r = finish_alt_dot(acc[qa * TILE + qb], P.metric);
                            if (r < M.thr[uo + a]) pm |= 1u << (2 * (qa * TILE + qb));
                            if (b != a && r < M.thr[uo + b]) pm |= 2u << (2 * (qa * TILE + qb));
                        }
                        acc[qa * TILE + qb] = r;
                    }
            }
            const int npush = __popc(pm);
            int incl = npush;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            unsigned long long wbase = 0;
            if (total > 0) {
                if (lane == 31) wbase = atomicAdd(P.cursor, (unsigned long long)total);
                wbase = __shfl_sync(0xffffffffu, wbase, 31);
            }
            if (npush > 0) {
                if (wbase + (unsigned long long)total > P.capacity) {
                    *P.overflow = 1;                        // the host discards the round; keep the pipeline handshakes going
                } else {
                    unsigned long long w = wbase + (unsigned long long)(incl - npush);
                    const unsigned long long unit = (unsigned long long)M.uid[u];
                    uint32_t mk = pm;
                    while (mk) {
                        const int bit = __ffs(mk) - 1; mk &= mk - 1;
                        const int pr = bit >> 1, side = bit & 1;
                        const int a = ta * TILE + pr / TILE, b = tb * TILE + pr % TILE;
                        const int32_t pa = M.ids[uo + a], pb = M.ids[uo + b];
                        const int32_t tgt = side ? pb : pa, oth = side ? pa : pb;
                        float r = 0.0f;
#pragma unroll
                        for (int sidx = 0; sidx < TILE * TILE; ++sidx) if (sidx == pr) r = acc[sidx];
                        P.out_tgt[w] = tgt; P.out_oth[w] = oth; P.out_d[w] = r;
                        // out_pos == nullptr: the rank inside the target's segment is taken by the scatter kernel, so this
                        // is a fire-and-forget reduction (RED) instead of an atomic whose result the store below waits for
                        if (P.out_pos) P.out_pos[w] = atomicAdd(&P.tgt_count[tgt], 1u);
                        else atomicAdd(&P.tgt_count[tgt], 1u);
                        P.out_seq[w] = (unit << 20) | ((unsigned long long)(a * P.slot_stride + b) << 1) | (unsigned long long)side;
                        ++w;
                    }
                }
            }
        }
        __syncwarp();                                       // every lane is done with the descriptor
        if (lane == 0) mbar_arrive(&mempty[q]);
    }
}

// ------------------------------------------------------------------------------------------------ host
size_t join_pipe_meta_bytes(int64_t n_batches) { return (size_t)(n_batches > 0 ? n_batches : 0) * JP_REC_BYTES + 256; }

template <int METRIC, int DKF, int NST>
static int launch_pipe_t(nnd_ctx *ctx, const JoinBatchParams &JB, const JpRec *recs, int ctas_per_sm, cudaStream_t st) {
    constexpr size_t smem = jp_smem_bytes<DKF, NST>();
    static_assert(smem <= 227 * 1024, "stage ring exceeds the shared memory of one CTA");
    NND_CUDA(cudaFuncSetAttribute(k_join_pipe<METRIC, DKF, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int fit = 0;
    NND_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fit, k_join_pipe<METRIC, DKF, NST>, JP_THREADS, smem));
    NND_CHECK(fit >= 1, NND_ERR_UNSUPPORTED, "k_join_pipe<%d,%d>: %zu B of shared memory do not fit", DKF, NST, smem);
    int per_sm = ctas_per_sm > 0 ? (ctas_per_sm < fit ? ctas_per_sm : fit) : fit;
    int64_t nb = JB.batch_end - JB.batch_begin, grid = (int64_t)ctx->sm_count * per_sm;
    if (grid > nb) grid = nb;
    k_join_pipe<METRIC, DKF, NST><<<(unsigned)grid, JP_THREADS, smem, st>>>(JB, recs);
    NND_LAUNCH(ctx);
    NND_CUDA(cudaGetLastError());
    return NND_OK;
}

template <int METRIC>
static int launch_pipe_m(nnd_ctx *ctx, const JoinBatchParams &JB, const JpRec *recs, const JoinPipeOpts &o, cudaStream_t st) {
    if (o.dk == 128) return launch_pipe_t<METRIC, 128, 2>(ctx, JB, recs, o.ctas_per_sm, st);   // whole 512-byte rows, 1 CTA/SM
    if (o.dk == 64) {
        if (o.stages <= 2) return launch_pipe_t<METRIC, 64, 2>(ctx, JB, recs, o.ctas_per_sm, st);
        if (o.stages == 3) return launch_pipe_t<METRIC, 64, 3>(ctx, JB, recs, o.ctas_per_sm, st);
        return launch_pipe_t<METRIC, 64, 4>(ctx, JB, recs, o.ctas_per_sm, st);
    }
    if (o.stages <= 2) return launch_pipe_t<METRIC, 32, 2>(ctx, JB, recs, o.ctas_per_sm, st);
    if (o.stages == 3) return launch_pipe_t<METRIC, 32, 3>(ctx, JB, recs, o.ctas_per_sm, st);
    if (o.stages <= 4) return launch_pipe_t<METRIC, 32, 4>(ctx, JB, recs, o.ctas_per_sm, st);
    return launch_pipe_t<METRIC, 32, 6>(ctx, JB, recs, o.ctas_per_sm, st);
}

int launch_join_pipe(nnd_ctx *ctx, const JoinBatchParams &JB, void *meta_buf, size_t meta_bytes, const JoinPipeOpts &opt,
                     cudaStream_t st) {
    const int64_t nb = JB.batch_end - JB.batch_begin;
    if (nb <= 0) return NND_OK;
    NND_CHECK(meta_buf && join_pipe_meta_bytes(nb) <= meta_bytes, NND_ERR_WORKSPACE,
              "k_join_pipe: %lld batches need %zu B of descriptor space (have %zu)", (long long)nb, join_pipe_meta_bytes(nb), meta_bytes);
    NND_CHECK(JB.ld % 4 == 0, NND_ERR_INVALID, "k_join_pipe: rows must be padded to 16 bytes");
    JpRec *recs = (JpRec *)meta_buf;
    const unsigned g = (unsigned)cdiv(nb, JPM_WARPS);
    if (JB.metric == NND_METRIC_ALT_COSINE) k_jp_meta<true><<<g, JPM_WARPS * 32, 0, st>>>(JB, recs);
    else k_jp_meta<false><<<g, JPM_WARPS * 32, 0, st>>>(JB, recs);
    NND_LAUNCH(ctx);
    NND_CUDA(cudaGetLastError());
    if (JB.metric == NND_METRIC_SQEUCLIDEAN) return launch_pipe_m<NND_METRIC_SQEUCLIDEAN>(ctx, JB, recs, opt, st);
    if (JB.metric == NND_METRIC_ALT_COSINE) return launch_pipe_m<NND_METRIC_ALT_COSINE>(ctx, JB, recs, opt, st);
    return launch_pipe_m<NND_METRIC_ALT_DOT>(ctx, JB, recs, opt, st);
}
