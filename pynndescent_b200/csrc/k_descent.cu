// k_descent.cu -- the NN-descent inner loop on sm_100a:
//   k_join    : all-pairs distances inside a leaf (init_rp_tree) or a candidate row (local join),
//               rows staged in shared memory by the bulk-copy (TMA) engine, 4x4 register tiles
//   k_scatter : counting-sort routing of half-updates to their target heap rows
//   k_apply   : one warp per heap row replays the pushes in the reference's canonical order
//
// Reference semantics (file:line under /root/reference/pynndescent):
//   generate_leaf_updates        pynndescent_.py:73-113   (strict '<' filter, pairs i<j)
//   generate_graph_update_array  utils.py:536-658         ('<=' filter, new x new (k>=j) then new x old)
//   apply_graph_update_array     utils.py:661-733         (push (d,q)->row p then (d,p)->row q, in emission order)
//   checked_flagged_heap_push    utils.py:471-533
//
// Exactness: a half-update (row t <- other o, dist dd) is emitted iff dd < thr[t] where thr is a
// snapshot of the heap roots.  The reference emits iff dd <= max(thr_p,thr_q) and then the push
// re-checks dd < root(now) <= thr, so every half-update dropped here would have been rejected,
// and the survivors are replayed per row in ascending emission sequence => identical heaps,
// identical flags, identical accepted-push count.
#include <stdlib.h>

#include <type_traits>
#include <vector>

#include "nnd_common.cuh"
#include "join_batch.cuh"

#define DK 128  // d-chunk staged per pass (floats)
#define SD (DK + 4)  // padded smem row stride: consecutive rows shift by 4 banks

struct JoinParams {
    const float *X; int64_t N; int d, ld; int metric; const void *aux;
    const float *thr;
    int mode;  // 0 local join, 1 leaf (routed), 2 leaf fused with apply, 3 leaf: distance matrix only (-> k_leaf_push)
    const int32_t *newc, *oldc; int C; int64_t unit_begin, unit_end;
    const int32_t *leaf_points; const int64_t *leaf_offsets;
    int m_max, slot_stride;
    int32_t *out_tgt, *out_oth; float *out_d; uint32_t *out_pos; unsigned long long *out_seq;
    uint32_t *tgt_count;
    unsigned long long *cursor; unsigned long long capacity;
    unsigned long long *stats;  // [0] rows gathered [1] pairs evaluated
    int *overflow;
    // mode 2 (fused leaf join + apply): heaps are updated in place by the leaf's own CTA
    int32_t *heap_idx; float *heap_dist; uint8_t *heap_flag; int k; int d_alias;
    int64_t row_lo, row_hi;   // mode 0: rows outside are skipped (sharded build)
    // mode 3 (split leaf init, NND_LEAF_SPLIT): the leaf's distance matrix goes to global memory, k_leaf_push applies it
    float *leaf_D; const uint32_t *leaf_doff;
};

// compile-time views of a metric CLASS (nnd_metric_class): per-row auxiliaries in shared memory, accumulator type
#define MC_AUX1(M) ((M) == NND_METRIC_ALT_COSINE || (M) == NND_METRIC_ALT_JACCARD || (M) == NND_METRIC_ALT_HELLINGER)
#define MC_AUX2(M) ((M) == NND_METRIC_CORRELATION)
#define MC_F64(M) ((M) == NND_METRIC_CORRELATION || (M) == NND_METRIC_MANHATTAN || (M) == NND_METRIC_CANBERRA || (M) == NND_METRIC_MINKOWSKI)

static size_t join_smem_bytes(int m_max, int metric) {
    size_t b = 16;                       // mbarrier
    b += (size_t)m_max * 4 * 2;          // ids, thr
    if (nnd_metric_aux_kind(metric) == 1) b += (size_t)m_max * 4;
    if (nnd_metric_aux_kind(metric) == 2) b += (size_t)m_max * 16;
    b = (b + 15) & ~size_t(15);
    b += (size_t)(m_max + 4) * SD * 4;  // slot permutation (j&3)*ntc + (j>>2) spans up to m+3 slots
    return b;
}
// fused leaf mode: + per-member heap columns [k][m_max] (f32,i32,u8) + distance matrix (aliased onto the
// row tiles when it fits)
static size_t fused_extra_bytes(int m_max, int k, int *alias) {
    size_t heap = ((size_t)k * m_max * 9 + 15) & ~size_t(15);
    int ds = m_max | 1;
    *alias = (ds <= SD);
    return heap + (*alias ? 0 : (size_t)m_max * ds * 4 + 16);
}

template <int METRIC, int TPT, int MAXT>
__global__ void __launch_bounds__(MAXT) k_join(JoinParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *bar = (uint64_t *)smem_raw;
    int32_t *ids = (int32_t *)(smem_raw + 16);
    float *thr_s = (float *)(ids + P.m_max);
    unsigned char *cur = (unsigned char *)(thr_s + P.m_max);
    float *n2_s = nullptr; double *mu_s = nullptr;
    if (MC_AUX1(METRIC)) { n2_s = (float *)cur; cur += (size_t)P.m_max * 4; }
    if (MC_AUX2(METRIC)) {
        cur = smem_align_up(smem_raw, cur, 8);
        mu_s = (double *)cur; cur += (size_t)P.m_max * 16;
    }
    float *hpri = nullptr; int32_t *hind = nullptr; uint8_t *hflg = nullptr; float *dsep = nullptr;
    if (P.mode == 2) {
        cur = smem_align_up(smem_raw, cur, 16);
        hpri = (float *)cur; hind = (int32_t *)(hpri + (size_t)P.k * P.m_max); hflg = (uint8_t *)(hind + (size_t)P.k * P.m_max);
        cur += ((size_t)P.k * P.m_max * 9 + 15) & ~size_t(15);
        if (!P.d_alias) { dsep = (float *)cur; cur += (size_t)P.m_max * (P.m_max | 1) * 4 + 16; }
    }
    float *rows = (float *)smem_align_up(smem_raw, cur, 16);
    __shared__ int s_nn, s_m;

    const int tid = threadIdx.x, lane = tid & 31, nthr = blockDim.x;
    const int64_t unit = P.unit_begin + blockIdx.x;
    if (unit >= P.unit_end) return;
    if (P.mode == 0 && (unit < P.row_lo || unit >= P.row_hi)) return;

    // ---- member list (order-preserving compaction of the -1 holes; warp 0) ----
    if (tid < 32) {
        int cnt = 0;
        if (P.mode == 0) {
            for (int part = 0; part < 2; ++part) {
                const int32_t *src = (part == 0 ? P.newc : P.oldc) + unit * P.C;
                for (int b0 = 0; b0 < P.C; b0 += 32) {
                    int j = b0 + lane;
                    int32_t v = (j < P.C) ? src[j] : -1;
                    unsigned mk = __ballot_sync(0xffffffffu, v >= 0);
                    if (v >= 0) ids[cnt + __popc(mk & ((1u << lane) - 1))] = v;
                    cnt += __popc(mk);
                }
                if (part == 0 && lane == 0) s_nn = cnt;
            }
        } else {
            int64_t o0 = P.leaf_offsets[unit], o1 = P.leaf_offsets[unit + 1];
            cnt = (int)(o1 - o0);
            for (int j = lane; j < cnt; j += 32) ids[j] = P.leaf_points[o0 + j];
            if (lane == 0) s_nn = cnt;
        }
        if (lane == 0) { s_m = cnt; mbar_init(bar, 1); mbar_fence_init(); }
    }
    __syncthreads();
    const int nn = s_nn, m = s_m;
    if (nn == 0 || (P.mode != 0 && m < 2)) return;
    for (int j = tid; j < m; j += nthr) {
        int32_t id = ids[j];
        thr_s[j] = (P.mode >= 2) ? CUDART_INF_F : P.thr[id];
        if (MC_AUX1(METRIC)) n2_s[j] = ((const float *)P.aux)[id];
        if (MC_AUX2(METRIC)) {
            mu_s[2 * j] = ((const double *)P.aux)[2 * (int64_t)id];
            mu_s[2 * j + 1] = ((const double *)P.aux)[2 * (int64_t)id + 1];
        }
    }
    if (P.mode == 2) {
        // stage the members' heap rows now (column layout [j][member]); the loads complete behind the
        // bulk copies and the distance phase
        const int k = P.k, hs = P.m_max;
        for (int idx = tid; idx < m * k; idx += nthr) {
            int a = idx / k, j = idx - a * k;
            const int64_t row = ids[a];
            hpri[j * hs + a] = P.heap_dist[row * k + j]; hind[j * hs + a] = P.heap_idx[row * k + j]; hflg[j * hs + a] = P.heap_flag[row * k + j];
        }
    }
    if (tid == 0 && P.stats) {
        atomicAdd(&P.stats[0], (unsigned long long)m);
        unsigned long long np = (P.mode == 0) ? ((unsigned long long)nn * (nn + 1) / 2 + (unsigned long long)nn * (m - nn))
                                              : ((unsigned long long)m * (m - 1) / 2);
        atomicAdd(&P.stats[1], np);
    }

    // ---- my tiles: (ta,tb) with tb >= ta over ceil(nn/4) x ceil(m/4) ----
    const int ntr = (nn + TILE - 1) / TILE, ntc = (m + TILE - 1) / TILE;
    const int ntiles = ntr * ntc - ntr * (ntr - 1) / 2;
    int my_ta[TPT], my_tb[TPT];
#pragma unroll
    for (int i = 0; i < TPT; ++i) {
        int t = tid + i * nthr;
        my_ta[i] = -1; my_tb[i] = 0;
        if (t < ntiles) {
            int ta = 0;
            while (t >= ntc - ta) { t -= ntc - ta; ++ta; }
            my_ta[i] = ta; my_tb[i] = ta + t;
        }
    }

    typedef typename std::conditional<MC_F64(METRIC), double, float>::type acc_t;
    const double mk_p = METRIC == NND_METRIC_MINKOWSKI ? ((const double *)P.aux)[0] : 0.0;   // distances.py:138 p (nnd_metric_set_param)
    acc_t acc[TPT][TILE * TILE];
#pragma unroll
    for (int i = 0; i < TPT; ++i)
#pragma unroll
        for (int p = 0; p < TILE * TILE; ++p) acc[i][p] = (acc_t)0;

    // ---- d-chunks: bulk-copy rows -> smem, then register-tiled sequential accumulation ----
    uint32_t phase = 0;
    for (int c0 = 0; c0 < P.ld; c0 += DK) {
        const int len = min(DK, P.ld - c0);
        if (tid < 32) {
            if (lane == 0) mbar_expect_tx(bar, (uint32_t)m * (uint32_t)len * 4u);
            __syncwarp();
            for (int r = lane; r < m; r += 32)
                bulk_g2s(rows + (size_t)((r & 3) * ntc + (r >> 2)) * SD, P.X + (int64_t)ids[r] * P.ld + c0, (uint32_t)len * 4u, bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
#pragma unroll
        for (int i = 0; i < TPT; ++i) {
            if (my_ta[i] < 0) continue;
            const float *ra[TILE], *rb[TILE];
            acc_t ma[TILE], mb[TILE];
#pragma unroll
            for (int u = 0; u < TILE; ++u) {
                int ia = min(my_ta[i] * TILE + u, nn - 1), ib = min(my_tb[i] * TILE + u, m - 1);
                // member j lives at smem slot (j&3)*ntc + (j>>2): the 8 lanes of an LDS.128 phase (consecutive
                // tile columns) then read consecutive slots = distinct bank groups (stride 132 floats)
                ra[u] = rows + (size_t)((ia & 3) * ntc + (ia >> 2)) * SD; rb[u] = rows + (size_t)((ib & 3) * ntc + (ib >> 2)) * SD;
                if (METRIC == NND_METRIC_CORRELATION) { ma[u] = (acc_t)mu_s[2 * ia]; mb[u] = (acc_t)mu_s[2 * ib]; }
                else { ma[u] = 0; mb[u] = 0; }
            }
            for (int e = 0; e < len; e += 4) {
                float4 va[TILE], vb[TILE];
#pragma unroll
                for (int u = 0; u < TILE; ++u) {
                    va[u] = *(const float4 *)(ra[u] + e);
                    vb[u] = *(const float4 *)(rb[u] + e);
                }
#pragma unroll
                for (int w = 0; w < 4; ++w) {
#pragma unroll
                    for (int u = 0; u < TILE; ++u) {
                        float xa = (w == 0) ? va[u].x : (w == 1) ? va[u].y : (w == 2) ? va[u].z : va[u].w;
#pragma unroll
                        for (int v = 0; v < TILE; ++v) {
                            float xb = (w == 0) ? vb[v].x : (w == 1) ? vb[v].y : (w == 2) ? vb[v].z : vb[v].w;
                            if (METRIC == NND_METRIC_SQEUCLIDEAN) {
                                float df = __fsub_rn(xa, xb);
                                acc[i][u * TILE + v] = __fmaf_rn(df, df, (float)acc[i][u * TILE + v]);
                            } else if (METRIC == NND_METRIC_CORRELATION) {
                                if (c0 + e + w < P.d) {
                                    double sx = __dadd_rn((double)xa, -(double)ma[u]);
                                    double sy = __dadd_rn((double)xb, -(double)mb[v]);
                                    acc[i][u * TILE + v] = __fma_rn(sx, sy, (double)acc[i][u * TILE + v]);
                                }
                            } else if (METRIC == NND_METRIC_ALT_COSINE || METRIC == NND_METRIC_ALT_DOT) {
                                acc[i][u * TILE + v] = __fmaf_rn(xa, xb, (float)acc[i][u * TILE + v]);
                            } else if (METRIC == NND_METRIC_MINKOWSKI) {
                                acc[i][u * TILE + v] = (acc_t)nnd_minkowski_step((double)acc[i][u * TILE + v], xa, xb, mk_p);
                            } else {
                                acc[i][u * TILE + v] = nnd_metric_step<METRIC, acc_t>(acc[i][u * TILE + v], xa, xb);
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();  // everyone done with this chunk before it is overwritten
    }

    // ---- finalise, filter, count pushes ----
    float dd[TPT][TILE * TILE];
    uint32_t pm[TPT];
    int npush = 0;
#pragma unroll
    for (int i = 0; i < TPT; ++i) {
        pm[i] = 0;
        if (my_ta[i] < 0) continue;
#pragma unroll
        for (int u = 0; u < TILE; ++u)
#pragma unroll
            for (int v = 0; v < TILE; ++v) {
                int a = my_ta[i] * TILE + u, b = my_tb[i] * TILE + v;
                bool valid = (a < nn) && (b < m) && ((b > a) || (b == a && P.mode == 0));
                float r = 0.0f;
                if (valid) {
                    if (METRIC == NND_METRIC_SQEUCLIDEAN) r = (float)acc[i][u * TILE + v];
                    else if (METRIC == NND_METRIC_ALT_COSINE) r = finish_alt_cosine((float)acc[i][u * TILE + v], n2_s[a], n2_s[b]);
                    else if (METRIC == NND_METRIC_ALT_DOT) r = finish_alt_dot((float)acc[i][u * TILE + v], P.metric);
                    else if (METRIC == NND_METRIC_CORRELATION) r = finish_correlation((double)acc[i][u * TILE + v], mu_s[2 * a + 1], mu_s[2 * b + 1]);
                    else if (METRIC == NND_METRIC_ALT_JACCARD) r = finish_binary(P.metric, (float)acc[i][u * TILE + v], n2_s[a], n2_s[b], P.d);
                    else if (METRIC == NND_METRIC_ALT_HELLINGER) r = finish_alt_hellinger((float)acc[i][u * TILE + v], n2_s[a], n2_s[b]);
                    else if (METRIC == NND_METRIC_CHEBYSHEV) r = (float)acc[i][u * TILE + v];
                    else if (METRIC == NND_METRIC_MINKOWSKI) r = finish_minkowski((double)acc[i][u * TILE + v], ((const double *)P.aux)[1]);
                    else r = finish_plain(METRIC, (double)acc[i][u * TILE + v], P.d);
                    if (r < thr_s[a]) pm[i] |= 1u << (2 * (u * TILE + v));
                    if (b != a && r < thr_s[b]) pm[i] |= 2u << (2 * (u * TILE + v));
                }
                dd[i][u * TILE + v] = r;
            }
        npush += __popc(pm[i]);
    }
    if (P.mode == 3) {
        // ---- split leaf init: the m x m distance matrix of the leaf -> global memory, both triangles, row b contiguous
        // over a, so that k_leaf_push's thread-per-member reads are coalesced
        float *Dg = P.leaf_D + P.leaf_doff[unit - P.unit_begin];
#pragma unroll
        for (int i = 0; i < TPT; ++i) {
            if (my_ta[i] < 0) continue;
#pragma unroll
            for (int u = 0; u < TILE; ++u)
#pragma unroll
                for (int v = 0; v < TILE; ++v) {
                    int a = my_ta[i] * TILE + u, b = my_tb[i] * TILE + v;
                    if (a < nn && b < m && b > a) { Dg[(size_t)a * m + b] = dd[i][u * TILE + v]; Dg[(size_t)b * m + a] = dd[i][u * TILE + v]; }
                }
        }
        return;
    }
    if (P.mode == 2) {
        // ---- fused apply: distance matrix -> smem, then ONE THREAD PER MEMBER replays its pushes.
        // Row ids[a] receives, in the reference's emission order (pairs (i,j), i<j, row-major), exactly
        // the other members in ascending member position: (i,a) for i<a, then (a,j) for j>a.
        float *Dm = P.d_alias ? rows : dsep;
        const int ds = m | 1;
#pragma unroll
        for (int i = 0; i < TPT; ++i) {
            if (my_ta[i] < 0) continue;
#pragma unroll
            for (int u = 0; u < TILE; ++u)
#pragma unroll
                for (int v = 0; v < TILE; ++v) {
                    int a = my_ta[i] * TILE + u, b = my_tb[i] * TILE + v;
                    if (a < nn && b < m && b > a) { Dm[a * ds + b] = dd[i][u * TILE + v]; Dm[b * ds + a] = dd[i][u * TILE + v]; }
                }
        }
        __syncthreads();
        const int hs = P.m_max, k = P.k;
        for (int a = tid; a < m; a += nthr) {
            const int64_t row = ids[a];
            float *pri = hpri + a; int32_t *ind = hind + a; uint8_t *flg = hflg + a;
            bool changed = false;
            for (int b = 0; b < m; ++b) {
                if (b == a) continue;
                const float p = Dm[a * ds + b];
                if (p >= pri[0]) continue;
                const int32_t nb = ids[b];
                bool dup = false;
                for (int j = 0; j < k; ++j) dup |= (ind[j * hs] == nb);
                if (dup) continue;
                int i = 0, i_swap;
                for (;;) {
                    int ic1 = 2 * i + 1, ic2 = ic1 + 1;
                    if (ic1 >= k) break;
                    else if (ic2 >= k) { if (pri[ic1 * hs] > p) i_swap = ic1; else break; }
                    else if (pri[ic1 * hs] >= pri[ic2 * hs]) { if (p < pri[ic1 * hs]) i_swap = ic1; else break; }
                    else { if (p < pri[ic2 * hs]) i_swap = ic2; else break; }
                    pri[i * hs] = pri[i_swap * hs]; ind[i * hs] = ind[i_swap * hs]; flg[i * hs] = flg[i_swap * hs];
                    i = i_swap;
                }
                pri[i * hs] = p; ind[i * hs] = nb; flg[i * hs] = 1;
                changed = true;
            }
            (void)row; (void)changed;
        }
        __syncthreads();
        for (int idx = tid; idx < m * k; idx += nthr) {   // coalesced write-back of all member rows
            int a = idx / k, j = idx - a * k;
            const int64_t row = ids[a];
            P.heap_dist[row * k + j] = hpri[j * hs + a]; P.heap_idx[row * k + j] = hind[j * hs + a]; P.heap_flag[row * k + j] = hflg[j * hs + a];
        }
        return;
    }
    // warp-aggregated slot allocation in the global push list
    int incl = npush;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    int total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned long long base = 0;
    if (total > 0) {
        if (lane == 31) base = atomicAdd(P.cursor, (unsigned long long)total);
        base = __shfl_sync(0xffffffffu, base, 31);
    }
    if (npush > 0) {
        unsigned long long w = base + (unsigned long long)(incl - npush);
        if (base + (unsigned long long)total > P.capacity) { *P.overflow = 1; return; }
#pragma unroll
        for (int i = 0; i < TPT; ++i) {
            uint32_t mk = pm[i];
            while (mk) {
                int bit = __ffs(mk) - 1;
                mk &= mk - 1;
                int pr = bit >> 1, side = bit & 1;
                int a = my_ta[i] * TILE + pr / TILE, b = my_tb[i] * TILE + pr % TILE;
                int32_t pa = ids[a], pb = ids[b];
                int32_t tgt = side ? pb : pa, oth = side ? pa : pb;
                // dd[i][pr] with a compile-time-unrollable select (keeps dd in registers)
                float r = 0.0f;
#pragma unroll
                for (int s = 0; s < TILE * TILE; ++s) if (s == pr) r = dd[i][s];
                P.out_tgt[w] = tgt; P.out_oth[w] = oth; P.out_d[w] = r;
                P.out_pos[w] = atomicAdd(&P.tgt_count[tgt], 1u);
                P.out_seq[w] = ((unsigned long long)unit << 20) |
                               ((unsigned long long)(a * P.slot_stride + b) << 1) | (unsigned long long)side;
                ++w;
            }
        }
    }
}


// =====================================================================================================
// k_join_batch: the local join for short rows (ld <= 256).  One 128-thread CTA handles a BATCH of consecutive
// candidate rows; every thread owns exactly one 4x4 tile of the batch's tile queue (batches are cut by a
// per-row weight = max(#member slots, 1.5 x #tiles) so that <= 192 slots and <= 128 tiles fit), so lanes stay
// busy even when one row has only a handful of tiles.  Member rows are staged by the TMA bulk-copy engine in
// d-chunks of 32 floats, DOUBLE-BUFFERED on two mbarriers: chunk c+1 is in flight while chunk c is accumulated
// (the sequential per-pair order is preserved because each thread walks the chunks in order).
// Same arithmetic and the same emission keys as k_join.
// =====================================================================================================
// prepass 1: per row -> active flag, physical slots, batching weight
__global__ void k_jb_count(const int32_t *newc, const int32_t *oldc, int C, int64_t N, uint32_t *flag, uint32_t *phys,
                           uint32_t *wgt, uint32_t *info, int64_t lo, int64_t hi) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const int64_t i = p;
    int nn = 0, no = 0;
    if (i >= lo && i < hi)
        for (int j = 0; j < C; ++j) { nn += newc[i * C + j] >= 0; no += oldc[i * C + j] >= 0; }
    int m = nn + no;
    uint32_t ph = nn > 0 ? (uint32_t)(((m + 3) >> 2) << 2) : 0u;
    uint32_t tw = nn > 0 ? (uint32_t)((jb_tiles(nn, m) * 3 + 1) >> 1) : 0u;
    flag[p] = nn > 0;
    phys[p] = ph;
    wgt[p] = ph > tw ? ph : tw;
    info[p] = (uint32_t)nn | ((uint32_t)m << 8);
}
// prepass 2: compacted active-row list + member ids in the reference's order (new k asc, then old k asc; holes dropped)
__global__ void k_jb_fill(const int32_t *newc, const int32_t *oldc, int C, int64_t N, const uint32_t *flag,
                          const uint32_t *pos, const uint32_t *poff, const uint32_t *woff, const uint32_t *info,
                          int32_t *act_unit, uint32_t *act_info, uint32_t *act_poff, uint32_t *act_woff, int32_t *members) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N || !flag[p]) return;
    const int64_t i = p;
    uint32_t j = pos[p], o = poff[p];
    act_unit[j] = (int32_t)i; act_info[j] = info[p]; act_poff[j] = o; act_woff[j] = woff[p];
    int w = 0;
    for (int c = 0; c < C; ++c) { int32_t v = newc[i * C + c]; if (v >= 0) members[o + w++] = v; }
    for (int c = 0; c < C; ++c) { int32_t v = oldc[i * C + c]; if (v >= 0) members[o + w++] = v; }
    int pe = ((w + 3) >> 2) << 2;
    for (; w < pe; ++w) members[o + w] = -1;
}
// prepass 3: first active row of every batch (batch b = rows whose weight prefix lies in [b*ps,(b+1)*ps))
__global__ void k_jb_bounds(const uint32_t *act_woff, int64_t n_active, uint32_t ps, int32_t *batch_first) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_active) return;
    uint32_t b = act_woff[j] / ps;
    if (j == 0 || act_woff[j - 1] / ps != b) batch_first[b] = (int32_t)j;
}

template <int METRIC, int JB_DK>
__global__ void __launch_bounds__(JB_THREADS) k_join_batch(JoinBatchParams P) {
    constexpr int JB_SD = JB_DK + 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *bar = (uint64_t *)smem_raw;                  // [2]
    int *s_hdr = (int *)(smem_raw + 16);                    // [0] nunits [1] tiles [2] members [3] total phys
    int32_t *s_uid = (int32_t *)(smem_raw + 32);            // [JB_MAXU]
    int *s_unn = s_uid + JB_MAXU, *s_um = s_unn + JB_MAXU;
    int *s_uoff = s_um + JB_MAXU;                           // [JB_MAXU+1]
    int *s_utile = s_uoff + JB_MAXU + 1;                    // [JB_MAXU+1]
    int32_t *ids = (int32_t *)(s_utile + JB_MAXU + 1);      // [JB_SCAP]
    float *thr_s = (float *)(ids + JB_SCAP);
    float *n2_s = thr_s + JB_SCAP;                          // cosine only
    float *rows = (float *)smem_align_up(smem_raw, (unsigned char *)(n2_s + (METRIC == NND_METRIC_ALT_COSINE ? JB_SCAP : 0)), 16);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t B = P.batch_begin + blockIdx.x;
    if (B >= P.batch_end) return;
    // batch_first[b] = first active row whose weight prefix starts in window b, or a sentinel (0x7f7f7f7f) when a
    // heavy row of the previous window covers window b entirely (then this CTA has nothing to do)
    int64_t j0 = P.batch_first[B];
    if (j0 >= P.n_active) return;
    int64_t j1 = P.n_active;
    for (int64_t bb = B + 1; bb < P.n_batches; ++bb) { int64_t v = P.batch_first[bb]; if (v < P.n_active) { j1 = v; break; } }
    const int nunits = (int)(j1 - j0);
    if (nunits <= 0) return;
    const uint32_t base = P.act_poff[j0];
    if (tid < nunits) {
        uint32_t inf = P.act_info[j0 + tid];
        s_uid[tid] = P.act_unit[j0 + tid]; s_unn[tid] = (int)(inf & 0xFF); s_um[tid] = (int)(inf >> 8);
        s_uoff[tid] = (int)(P.act_poff[j0 + tid] - base);
    }
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
    __syncthreads();
    if (tid == 0) {
        int t = 0, mem = 0;
        unsigned long long pairs = 0;
        for (int u = 0; u < nunits; ++u) {
            s_utile[u] = t;
            int nn = s_unn[u], m = s_um[u];
            t += jb_tiles(nn, m);
            mem += m;
            pairs += (unsigned long long)nn * (nn + 1) / 2 + (unsigned long long)nn * (m - nn);
        }
        s_utile[nunits] = t;
        s_uoff[nunits] = s_uoff[nunits - 1] + (((s_um[nunits - 1] + 3) >> 2) << 2);
        s_hdr[1] = t; s_hdr[2] = mem;
        if (P.stats) { atomicAdd(&P.stats[0], (unsigned long long)mem); atomicAdd(&P.stats[1], pairs); }
    }
    __syncthreads();
    const int totphys = s_uoff[nunits], mem = s_hdr[2], T = s_hdr[1];
    for (int e = tid; e < totphys; e += JB_THREADS) {
        int32_t id = P.members[base + e];
        ids[e] = id;
        thr_s[e] = id >= 0 ? P.thr[id] : 0.0f;
        if (METRIC == NND_METRIC_ALT_COSINE) n2_s[e] = id >= 0 ? ((const float *)P.aux)[id] : 0.0f;
    }
    // ---- my tile
    int u = -1, ta = 0, tb = 0, nn = 0, m = 0, uo = 0, ntc = 1;
    if (tid < T) {
        int lo = 0, hi = nunits;                        // last u with s_utile[u] <= tid
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (s_utile[mid] <= tid) lo = mid; else hi = mid; }
        u = lo;
        nn = s_unn[u]; m = s_um[u]; uo = s_uoff[u]; ntc = (m + 3) >> 2;
        int lt = tid - s_utile[u];
        while (lt >= ntc - ta) { lt -= ntc - ta; ++ta; }
        tb = ta + lt;
    }
    int oa[TILE], ob[TILE];                                 // smem row offsets (floats) of my 4 + 4 member rows
#pragma unroll
    for (int q = 0; q < TILE; ++q) {
        int ia = min(ta * TILE + q, max(nn - 1, 0)), ib = min(tb * TILE + q, max(m - 1, 0));
        oa[q] = (uo + (ia & 3) * ntc + (ia >> 2)) * JB_SD;
        ob[q] = (uo + (ib & 3) * ntc + (ib >> 2)) * JB_SD;
    }
    float acc[TILE * TILE];
#pragma unroll
    for (int p = 0; p < TILE * TILE; ++p) acc[p] = 0.0f;

    // ---- chunk pipeline: gather chunk c+1 while accumulating chunk c
    const int nchunks = (P.ld + JB_DK - 1) / JB_DK;
    auto expect = [&](int c) {
        if (tid == 0) mbar_expect_tx(&bar[c & 1], (uint32_t)mem * (uint32_t)min(JB_DK, P.ld - c * JB_DK) * 4u);
    };
    auto issue = [&](int c) {   // caller guarantees: expect(c) done and buffer c&1 free, both ordered by a __syncthreads
        const int c0 = c * JB_DK, len = min(JB_DK, P.ld - c0);
        float *buf = rows + (size_t)(c & 1) * JB_SCAP * JB_SD;
        for (int uu = warp; uu < nunits; uu += JB_THREADS / 32) {
            const int mm = s_um[uu], nt = (mm + 3) >> 2, o = s_uoff[uu];
            for (int j = lane; j < mm; j += 32)
                bulk_g2s(buf + (size_t)(o + (j & 3) * nt + (j >> 2)) * JB_SD, P.X + (int64_t)ids[o + j] * P.ld + c0,
                         (uint32_t)len * 4u, &bar[c & 1]);
        }
    };
    expect(0);
    if (nchunks > 1) expect(1);
    __syncthreads();                                        // ids[] complete, expect_tx posted
    issue(0);
    if (nchunks > 1) issue(1);
    for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&bar[c & 1], (uint32_t)((c >> 1) & 1));
        if (u >= 0) {
            const float *buf = rows + (size_t)(c & 1) * JB_SCAP * JB_SD;
            const int len = min(JB_DK, P.ld - c * JB_DK);
#pragma unroll 2
            for (int e = 0; e < len; e += 4) {
                float4 va[TILE], vb[TILE];
#pragma unroll
                for (int q = 0; q < TILE; ++q) { va[q] = *(const float4 *)(buf + oa[q] + e); vb[q] = *(const float4 *)(buf + ob[q] + e); }
#pragma unroll
                for (int w = 0; w < 4; ++w) {
#pragma unroll
                    for (int qa = 0; qa < TILE; ++qa) {
                        float xa = (w == 0) ? va[qa].x : (w == 1) ? va[qa].y : (w == 2) ? va[qa].z : va[qa].w;
#pragma unroll
                        for (int qb = 0; qb < TILE; ++qb) {
                            float xb = (w == 0) ? vb[qb].x : (w == 1) ? vb[qb].y : (w == 2) ? vb[qb].z : vb[qb].w;
                            if (METRIC == NND_METRIC_SQEUCLIDEAN) {
                                float df = __fsub_rn(xa, xb);
                                acc[qa * TILE + qb] = __fmaf_rn(df, df, acc[qa * TILE + qb]);
                            } else {
                                acc[qa * TILE + qb] = __fmaf_rn(xa, xb, acc[qa * TILE + qb]);
                            }
                        }
                    }
                }
            }
        }
        if (c + 2 < nchunks) expect(c + 2);                 // phase (c+2)>>1 of barrier c&1 (its previous phase completed: we waited on it)
        __syncthreads();                                    // everyone done with buffer c&1
        if (c + 2 < nchunks) issue(c + 2);
    }
    // ---- finalise + filter
    uint32_t pm = 0;
    if (u >= 0) {
#pragma unroll
        for (int qa = 0; qa < TILE; ++qa)
#pragma unroll
            for (int qb = 0; qb < TILE; ++qb) {
                int a = ta * TILE + qa, b = tb * TILE + qb;
                bool valid = (a < nn) && (b < m) && (b >= a);
                float r = 0.0f;
                if (valid) {
                    if (METRIC == NND_METRIC_SQEUCLIDEAN) r = acc[qa * TILE + qb];
                    else if (METRIC == NND_METRIC_ALT_COSINE) r = finish_alt_cosine(acc[qa * TILE + qb], n2_s[uo + a], n2_s[uo + b]);
                    else r = finish_alt_dot(acc[qa * TILE + qb], P.metric);
                    if (r < thr_s[uo + a]) pm |= 1u << (2 * (qa * TILE + qb));
                    if (b != a && r < thr_s[uo + b]) pm |= 2u << (2 * (qa * TILE + qb));
                }
                acc[qa * TILE + qb] = r;
            }
    }
    int npush = __popc(pm), incl = npush;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    int total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned long long wbase = 0;
    if (total > 0) {
        if (lane == 31) wbase = atomicAdd(P.cursor, (unsigned long long)total);
        wbase = __shfl_sync(0xffffffffu, wbase, 31);
    }
    if (npush > 0) {
        if (wbase + (unsigned long long)total > P.capacity) { *P.overflow = 1; return; }
        unsigned long long w = wbase + (unsigned long long)(incl - npush);
        const unsigned long long unit = (unsigned long long)s_uid[u];
        uint32_t mk = pm;
        while (mk) {
            int bit = __ffs(mk) - 1; mk &= mk - 1;
            int pr = bit >> 1, side = bit & 1;
            int a = ta * TILE + pr / TILE, b = tb * TILE + pr % TILE;
            int32_t pa = ids[uo + a], pb = ids[uo + b];
            int32_t tgt = side ? pb : pa, oth = side ? pa : pb;
            float r = 0.0f;
#pragma unroll
            for (int sidx = 0; sidx < TILE * TILE; ++sidx) if (sidx == pr) r = acc[sidx];
            P.out_tgt[w] = tgt; P.out_oth[w] = oth; P.out_d[w] = r;
            // out_pos == nullptr (NND_RANK_SCATTER): k_scatter_rank takes the rank inside the target's segment, so this is
            // a fire-and-forget reduction (RED) instead of an atomic whose result the store has to wait for
            if (P.out_pos) P.out_pos[w] = atomicAdd(&P.tgt_count[tgt], 1u);
            else atomicAdd(&P.tgt_count[tgt], 1u);
            P.out_seq[w] = (unit << 20) | ((unsigned long long)(a * P.slot_stride + b) << 1) | (unsigned long long)side;
            ++w;
        }
    }
}

// ------------------------------------------------------------------ routing
__global__ void k_snapshot_thr(const float *heap_dist, int k, float *thr, int64_t N) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) thr[i] = heap_dist[i * k];
}

__global__ void k_scatter(const int32_t *tgt, const int32_t *oth, const float *dd, const uint32_t *pos,
                          const unsigned long long *seq, unsigned long long n, const uint32_t *off, Rec16 *rec) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Rec16 r; r.key = seq[i]; r.oth = oth[i]; r.d = dd[i];
    rec[(size_t)off[tgt[i]] + pos[i]] = r;
}

// NND_RANK_SCATTER: the join only counted its targets; the rank inside the target's segment is taken here, where every
// thread has one independent record in flight (segment order is arbitrary either way: k_apply sorts by sequence)
__global__ void k_scatter_rank(const int32_t *tgt, const int32_t *oth, const float *dd, const unsigned long long *seq,
                               unsigned long long n, const uint32_t *off, uint32_t *rank, Rec16 *rec) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t t = tgt[i];
    Rec16 r; r.key = seq[i]; r.oth = oth[i]; r.d = dd[i];
    rec[(size_t)off[t] + atomicAdd(&rank[t], 1u)] = r;
}

// ------------------------------------------------------------------ apply
#define APPLY_WARPS 8
__global__ void __launch_bounds__(APPLY_WARPS * 32) k_apply(int32_t *heap_idx, float *heap_dist, uint8_t *heap_flag,
                                                            int k, int64_t N, const uint32_t *cnt, const uint32_t *off,
                                                            Rec16 *rec, unsigned long long *c_total, int skip_small,
                                                            int short_sort) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * APPLY_WARPS + warp;
    if (row >= N) return;
    const uint32_t n = cnt[row];
    if (n == 0 || (skip_small && n <= 48)) return;
    Rec16 *buf = (Rec16 *)smem_raw + (size_t)warp * NND_SEG_SMEM;
    unsigned char *hp = smem_raw + (size_t)APPLY_WARPS * NND_SEG_SMEM * sizeof(Rec16) + (size_t)warp * (((size_t)k * 9 + 15) & ~size_t(15));
    float *pri = (float *)hp; int32_t *ind = (int32_t *)(pri + k); uint8_t *flg = (uint8_t *)(ind + k);
    for (int j = lane; j < k; j += 32) {
        pri[j] = heap_dist[row * k + j]; ind[j] = heap_idx[row * k + j]; flg[j] = heap_flag[row * k + j];
    }
    __syncwarp();
    const float root0 = pri[0];
    Rec16 *seg = rec + off[row];
    Rec16 *sorted; int ns;
    if (n <= NND_SEG_SMEM) {
        // keep only pushes that can possibly be accepted (root only decreases), compact into smem
        int kept = 0;
        for (uint32_t b0 = 0; b0 < n; b0 += 32) {
            uint32_t j = b0 + lane;
            Rec16 r; r.key = ~0ULL; r.oth = -1; r.d = 0.f;
            bool keep = false;
            if (j < n) { r = seg[j]; keep = r.d < root0; }
            unsigned mk = __ballot_sync(0xffffffffu, keep);
            if (keep) buf[kept + __popc(mk & ((1u << lane) - 1))] = r;
            kept += __popc(mk);
        }
        if (kept == 0) return;
        __syncwarp();
        if (kept <= 32) {
            // common case: sort (key, slot) pairs in registers with shuffles -- no shared-memory traffic
            unsigned long long key = lane < kept ? buf[lane].key : ~0ULL;
            int slot = lane;
            // NND_APPLY_SHORT_SORT: a bitonic network over the first next_pow2(kept) lanes sorts them just as well (the other
            // lanes hold the maximum key and never mix with them) -- 1-3 stages instead of 15 for the 1-3 records a touched
            // row typically receives after the first iterations
            const int top = (short_sort & 1) ? (int)next_pow2_u32((uint32_t)(kept > 2 ? kept : 2)) : 32;
#pragma unroll
            for (int k2 = 2; k2 <= 32; k2 <<= 1) {
                if (k2 > top) break;
#pragma unroll
                for (int j = k2 >> 1; j > 0; j >>= 1) {
                    unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, j);
                    int os = __shfl_xor_sync(0xffffffffu, slot, j);
                    bool up = ((lane & k2) == 0), lower = ((lane & j) == 0);
                    bool take = (lower == up) ? (ok < key) : (ok > key);
                    if (take) { key = ok; slot = os; }
                }
            }
            // lane s now owns the s-th record in emission order.  Two filters that cannot change the row (see JoinBatchParams):
            // a source that already sits in the row as loaded is rejected whenever its turn comes, and of several records with
            // the same source only the first in emission order can ever be accepted -- both are dropped before the replay.
            Rec16 mine; mine.key = ~0ULL; mine.oth = -2 - lane; mine.d = 0.f;
            if (lane < kept) mine = buf[slot];
            bool live = lane < kept;
            if (short_sort & 2) {       // (only while every distance in the row is the metric's own: nnd_set_heap_trusted)
                for (int j = 0; j < k; ++j) live = live && (ind[j] != mine.oth);
                const unsigned grp = __match_any_sync(0xffffffffu, live ? mine.oth : -2 - lane);
                live = live && (lane == __ffs(grp) - 1);
            }
            unsigned lm = __ballot_sync(0xffffffffu, live);
            int c = 0;
            while (lm) {
                const int s2 = __ffs(lm) - 1; lm &= lm - 1;
                const float dd = __shfl_sync(0xffffffffu, mine.d, s2);
                const int32_t oo = __shfl_sync(0xffffffffu, mine.oth, s2);
                c += warp_heap_push(pri, ind, flg, k, dd, oo, (uint8_t)1, lane);
            }
            if (c > 0) {
                for (int j = lane; j < k; j += 32) {
                    heap_dist[row * k + j] = pri[j]; heap_idx[row * k + j] = ind[j]; heap_flag[row * k + j] = flg[j];
                }
                if (lane == 0) atomicAdd(c_total, (unsigned long long)c);
            }
            return;
        }
        int p2 = (int)next_pow2_u32((uint32_t)kept);
        for (int j = kept + lane; j < p2; j += 32) { buf[j].key = ~0ULL; buf[j].oth = -1; buf[j].d = CUDART_INF_F; }
        __syncwarp();
        warp_bitonic_sort_rec(buf, p2, lane);
        sorted = buf; ns = kept;
    } else {
        uint32_t cap = padded_seg(n);
        for (uint32_t j = n + lane; j < cap; j += 32) { seg[j].key = ~0ULL; seg[j].oth = -1; seg[j].d = CUDART_INF_F; }
        __syncwarp();
        warp_bitonic_sort_rec(seg, (int)cap, lane);
        sorted = seg; ns = (int)n;
    }
    int c = 0;
    for (int s = 0; s < ns; ++s) {
        Rec16 r = sorted[s];
        c += warp_heap_push(pri, ind, flg, k, r.d, r.oth, (uint8_t)1, lane);
    }
    if (c > 0) {
        for (int j = lane; j < k; j += 32) {
            heap_dist[row * k + j] = pri[j]; heap_idx[row * k + j] = ind[j]; heap_flag[row * k + j] = flg[j];
        }
        if (lane == 0) atomicAdd(c_total, (unsigned long long)c);
    }
}


__global__ void k_row_bound(const int32_t *newc, const int32_t *oldc, int C, int64_t N, uint32_t *bound, int64_t lo, int64_t hi) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int nn = 0, no = 0;
    if (i >= lo && i < hi)
    for (int j = 0; j < C; ++j) { nn += newc[i * C + j] >= 0; no += oldc[i * C + j] >= 0; }
    bound[i] = (uint32_t)(nn * (nn + 1) + 2 * nn * no);  // 2 half-updates per pair (self pairs: 1, bounded by 2)
}
__global__ void k_leaf_bound(const int64_t *leaf_offsets, int64_t n_leaves, uint32_t *bound) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_leaves) return;
    uint32_t m = (uint32_t)(leaf_offsets[i + 1] - leaf_offsets[i]);
    bound[i] = m * (m - (m > 0));
}

// ------------------------------------------------------------------ split leaf init (NND_LEAF_SPLIT, default off)
// The fused leaf kernel keeps a 42 KB row buffer per CTA alive while ONE THREAD PER MEMBER replays the pushes: 4 CTAs per SM,
// two busy warps each, long dependent shared-memory chains (ncu_leaf_r01.txt: 16 % of the warp slots active, issue 29 %).
// Split: k_join mode 3 only produces the leaf's distance matrix (global, 4 m^2 bytes per leaf), and k_leaf_push replays the
// pushes with nothing but the heap columns in shared memory (9k bytes per member): 1.5k resident threads per SM instead of
// ~170.  Same pushes in the same order (the other members in ascending leaf position), so the heaps are identical.
__global__ void k_leaf_prep(const int64_t *leaf_offsets, int64_t l0, int64_t l1, int64_t pos0, uint32_t *dsz, int32_t *pos2leaf) {
    int64_t l = l0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= l1) return;
    const int64_t o0 = leaf_offsets[l], o1 = leaf_offsets[l + 1];
    const uint32_t m = (uint32_t)(o1 - o0);
    dsz[l - l0] = m * m;
    for (int64_t p = o0; p < o1; ++p) pos2leaf[p - pos0] = (int32_t)(l - l0);
}
#define LP_T 128
__global__ void __launch_bounds__(LP_T) k_leaf_push(const int32_t *leaf_points, const int64_t *leaf_offsets, int64_t l0, int64_t pos0,
                                                    int64_t npos, const int32_t *pos2leaf, const uint32_t *doff, const float *D,
                                                    int32_t *heap_idx, float *heap_dist, uint8_t *heap_flag, int k) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    float *hpri = (float *)smem_raw;                                   // [k][LP_T]
    int32_t *hind = (int32_t *)(hpri + (size_t)k * LP_T);
    uint8_t *hflg = (uint8_t *)(hind + (size_t)k * LP_T);
    int32_t *s_row = (int32_t *)(smem_raw + (((size_t)k * LP_T * 9 + 15) & ~size_t(15)));   // [LP_T] heap row of every thread's member
    const int64_t p0 = (int64_t)blockIdx.x * LP_T, p = p0 + tid;
    const int nth = (int)min((int64_t)LP_T, npos - p0);
    s_row[tid] = p < npos ? leaf_points[pos0 + p] : -1;
    __syncthreads();
    for (int e = tid; e < nth * k; e += LP_T) {            // the tile's heap rows (a gather by point id), 15 lanes per row
        const int r = e / k, j = e - r * k; const int64_t row = s_row[r];
        hpri[j * LP_T + r] = heap_dist[row * k + j]; hind[j * LP_T + r] = heap_idx[row * k + j]; hflg[j * LP_T + r] = heap_flag[row * k + j];
    }
    __syncthreads();
    if (p < npos) {
        const int l = pos2leaf[p];
        const int64_t o0 = leaf_offsets[l0 + l];
        const int m = (int)(leaf_offsets[l0 + l + 1] - o0), a = (int)(pos0 + p - o0);
        const float *Dl = D + doff[l];
        float *pri = hpri + tid; int32_t *ind = hind + tid; uint8_t *flg = hflg + tid;
        for (int b = 0; b < m; ++b) {
            if (b == a) continue;
            const float pr = Dl[(size_t)b * m + a];
            if (pr >= pri[0]) continue;
            const int32_t nb = leaf_points[o0 + b];
            bool dup = false;
            for (int j = 0; j < k; ++j) dup |= (ind[j * LP_T] == nb);
            if (dup) continue;
            int i = 0, i_swap;
            for (;;) {
                int ic1 = 2 * i + 1, ic2 = ic1 + 1;
                if (ic1 >= k) break;
                else if (ic2 >= k) { if (pri[ic1 * LP_T] > pr) i_swap = ic1; else break; }
                else if (pri[ic1 * LP_T] >= pri[ic2 * LP_T]) { if (pr < pri[ic1 * LP_T]) i_swap = ic1; else break; }
                else { if (pr < pri[ic2 * LP_T]) i_swap = ic2; else break; }
                pri[i * LP_T] = pri[i_swap * LP_T]; ind[i * LP_T] = ind[i_swap * LP_T]; flg[i * LP_T] = flg[i_swap * LP_T];
                i = i_swap;
            }
            pri[i * LP_T] = pr; ind[i * LP_T] = nb; flg[i * LP_T] = 1;
        }
    }
    __syncthreads();
    for (int e = tid; e < nth * k; e += LP_T) {
        const int r = e / k, j = e - r * k; const int64_t row = s_row[r];
        heap_dist[row * k + j] = hpri[j * LP_T + r]; heap_idx[row * k + j] = hind[j * LP_T + r]; heap_flag[row * k + j] = hflg[j * LP_T + r];
    }
}

// ------------------------------------------------------------------ host orchestration
struct DescentWs {
    float *thr; uint32_t *cnt, *off, *bound, *bound_scan, *scan_tmp;
    int32_t *l_tgt, *l_oth; float *l_d; uint32_t *l_pos; unsigned long long *l_seq;
    Rec16 *rec;
    unsigned long long *dev_scalars;  // [0] cursor [1] c_total [2..3] stats [4] overflow(int)
    // candidate-build scratch lives behind these (k_cand.cu)
    char *cand_base; size_t cand_bytes;
    int32_t *newc, *oldc;
    // batched local join prepass
    uint32_t *jb_flag, *jb_phys, *jb_wgt, *jb_info, *jb_pos, *jb_poff, *jb_woff, *act_info, *act_poff, *act_woff;
    int32_t *act_unit, *members, *batch_first;
};

size_t nnd_cand_ws_bytes(int64_t N, int k, int C, int T);

static int carve_descent_ws(void *ws, size_t ws_bytes, int64_t N, int k, int C, int T, int64_t cap, DescentWs *W,
                            size_t *need) {
    WsAlloc A(ws, ws_bytes);
    int64_t Nb = N > 16 ? N : 16;
    W->thr = A.get<float>(Nb);
    W->cnt = A.get<uint32_t>(Nb + 1);
    W->off = A.get<uint32_t>(Nb + 2);
    W->bound = A.get<uint32_t>(Nb + 1);
    W->bound_scan = A.get<uint32_t>(Nb + 2);
    W->scan_tmp = A.get<uint32_t>(nnd_scan_tmp_elems(Nb + 2));
    W->l_tgt = A.get<int32_t>(cap); W->l_oth = A.get<int32_t>(cap); W->l_d = A.get<float>(cap);
    W->l_pos = A.get<uint32_t>(cap); W->l_seq = A.get<unsigned long long>(cap);
    W->rec = A.get<Rec16>((size_t)cap * 2 + NND_SEG_SMEM);
    W->dev_scalars = A.get<unsigned long long>(16);
    // everything the stage calls share must sit BEFORE the T-dependent candidate scratch: nnd_descent carves
    // with the real T, the stage entry points with T=1, and both must agree on these offsets
    W->jb_flag = A.get<uint32_t>(Nb + 1); W->jb_phys = A.get<uint32_t>(Nb + 1); W->jb_info = A.get<uint32_t>(Nb + 1);
    W->jb_pos = A.get<uint32_t>(Nb + 2); W->jb_poff = A.get<uint32_t>(Nb + 2);
    W->jb_wgt = A.get<uint32_t>(Nb + 1); W->jb_woff = A.get<uint32_t>(Nb + 2);
    W->act_info = A.get<uint32_t>(Nb + 1); W->act_poff = A.get<uint32_t>(Nb + 2); W->act_unit = A.get<int32_t>(Nb + 1);
    W->act_woff = A.get<uint32_t>(Nb + 2);
    {   // one entry per weight window: N * maxw / (JB_SCAP - maxw) windows at most (depends on C, not on T)
        int Cc = C < 1 ? 1 : (C > JB_CMAX ? JB_CMAX : C);
        int mp = ((2 * Cc + 3) >> 2) << 2, mt = (jb_tiles(Cc, 2 * Cc) * 3 + 1) >> 1, mw = mp > mt ? mp : mt;
        int psw = JB_SCAP - mw; if (psw < 1) psw = 1;
        W->batch_first = A.get<int32_t>((size_t)Nb * (size_t)(mw / psw + 1) + 8);
    }
    W->members = A.get<int32_t>((size_t)Nb * (2 * JB_CMAX + 4));
    W->cand_bytes = nnd_cand_ws_bytes(N, k, C, T);
    W->cand_base = A.get<char>(W->cand_bytes);
    W->newc = A.get<int32_t>((size_t)Nb * C); W->oldc = A.get<int32_t>((size_t)Nb * C);
    if (need) *need = A.off;
    return A.ok ? NND_OK : NND_ERR_WORKSPACE;
}

extern "C" int nnd_descent_workspace_bytes(int64_t N, int32_t k, int32_t C, int32_t T, int64_t push_capacity,
                                           size_t *bytes) {
    NND_CHECK(bytes && N > 0 && k > 0 && push_capacity > 0, NND_ERR_INVALID, "nnd_descent_workspace_bytes: bad args");
    DescentWs W; size_t need = 0;
    carve_descent_ws(nullptr, 0, N, k, C, T, push_capacity, &W, &need);
    *bytes = need + 4096;
    return NND_OK;
}

template <int METRIC>
static int launch_join_m(nnd_ctx *ctx, const JoinParams &P, int64_t n_units, cudaStream_t st) {
    // worst-case tile count for this launch decides (TPT, threads)
    int nn_max = (P.mode == 0) ? P.C : P.m_max;
    int ntr = (nn_max + TILE - 1) / TILE, ntc = (P.m_max + TILE - 1) / TILE;
    int ntiles = ntr * ntc - ntr * (ntr - 1) / 2;
    size_t smem = join_smem_bytes(P.m_max, METRIC);
    if (P.mode == 2) { int al; smem += fused_extra_bytes(P.m_max, P.k, &al) + 32; }
    NND_CHECK(smem <= 220 * 1024, NND_ERR_UNSUPPORTED, "join: %d members need %zu B of shared memory", P.m_max, smem);
    int tpt = MC_F64(METRIC) ? 1 : (ntiles <= 512 ? 1 : (ntiles <= 2048 ? 2 : 4));
    // leaves are on average ~60 % full: two tiles per thread keep more lanes busy than one thread per tile slot
    static int leaf_tpt2 = -1;
    if (leaf_tpt2 < 0) { const char *e = getenv("NND_LEAF_TPT2"); leaf_tpt2 = e ? atoi(e) : 0; }
    if (P.mode == 2 && tpt == 1 && leaf_tpt2 && !MC_F64(METRIC) && ntiles > 64) tpt = 2;
    int threads = (int)(cdiv(cdiv(ntiles, tpt), 32) * 32);
    NND_CHECK(threads <= 1024, NND_ERR_UNSUPPORTED, "join: %d tiles exceed one CTA (members=%d)", ntiles, P.m_max);
    if (threads < 32) threads = 32;
    dim3 grid((unsigned)n_units);
#define LJ(T_, M_)                                                                                       \
    do {                                                                                                 \
        NND_CUDA(cudaFuncSetAttribute(k_join<METRIC, T_, M_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_join<METRIC, T_, M_><<<grid, threads, smem, st>>>(P);                                          \
    } while (0)
    if (tpt == 1) { if (threads <= 128) LJ(1, 128); else if (threads <= 512) LJ(1, 512); else LJ(1, 1024); }
    else if constexpr (!MC_F64(METRIC)) {           // f64 classes always run one tile per thread
        if (tpt == 2) { if (threads <= 512) LJ(2, 512); else LJ(2, 1024); }
        else LJ(4, 1024);
    }
#undef LJ
    NND_LAUNCH(ctx);
    NND_CUDA(cudaGetLastError());
    return NND_OK;
}
static int launch_join(nnd_ctx *ctx, const JoinParams &P, int64_t n_units, cudaStream_t st) {
    NND_CHECK(n_units < (1LL << 31), NND_ERR_INVALID, "join: too many units in one launch");
    switch (P.metric) {
    case NND_METRIC_SQEUCLIDEAN: return launch_join_m<NND_METRIC_SQEUCLIDEAN>(ctx, P, n_units, st);
    case NND_METRIC_ALT_COSINE: return launch_join_m<NND_METRIC_ALT_COSINE>(ctx, P, n_units, st);
    case NND_METRIC_ALT_DOT:
    case NND_METRIC_ALT_INNER: return launch_join_m<NND_METRIC_ALT_DOT>(ctx, P, n_units, st);   // one class, P.metric picks the tail
    case NND_METRIC_CORRELATION: return launch_join_m<NND_METRIC_CORRELATION>(ctx, P, n_units, st);
    case NND_METRIC_MANHATTAN: return launch_join_m<NND_METRIC_MANHATTAN>(ctx, P, n_units, st);
    case NND_METRIC_CHEBYSHEV: return launch_join_m<NND_METRIC_CHEBYSHEV>(ctx, P, n_units, st);
    case NND_METRIC_CANBERRA: return launch_join_m<NND_METRIC_CANBERRA>(ctx, P, n_units, st);
    case NND_METRIC_HAMMING: return launch_join_m<NND_METRIC_HAMMING>(ctx, P, n_units, st);
    case NND_METRIC_ALT_HELLINGER: return launch_join_m<NND_METRIC_ALT_HELLINGER>(ctx, P, n_units, st);
    case NND_METRIC_MINKOWSKI: return launch_join_m<NND_METRIC_MINKOWSKI>(ctx, P, n_units, st);
    }
    if (nnd_metric_class(P.metric) == NND_METRIC_ALT_JACCARD) return launch_join_m<NND_METRIC_ALT_JACCARD>(ctx, P, n_units, st);
    nnd_set_error("join: unknown metric %d", P.metric);
    return NND_ERR_UNSUPPORTED;
}

// one routing round: join over units [u0,u1) -> scan -> scatter -> apply.  Returns pushes routed.
static int launch_join_batch(nnd_ctx *ctx, JoinBatchParams &JB, cudaStream_t st);
static int route_round(nnd_ctx *ctx, const nnd_problem *P, JoinParams &J, int64_t u0, int64_t u1, DescentWs &W,
                       int64_t cap, unsigned long long *n_pushed, cudaStream_t st, JoinBatchParams *JB = nullptr,
                       bool do_apply = true) {
    const int64_t N = P->N;
    NND_STAGE(ctx, NND_ST_SCATTER, st);
    // sharded build: the snapshot of all N heap roots comes from the caller (all-gathered from the owners, nnd_set_thresholds)
    const float *thr = ctx->thr_ext ? ctx->thr_ext : W.thr;
    if (!ctx->thr_ext) { k_snapshot_thr<<<(unsigned)cdiv(N, 256), 256, 0, st>>>(P->heap_dist, P->k, W.thr, N); NND_LAUNCH(ctx); }
    NND_CUDA(cudaMemsetAsync(W.cnt, 0, sizeof(uint32_t) * (N + 1), st));
    NND_CUDA(cudaMemsetAsync(W.dev_scalars, 0, sizeof(unsigned long long), st));  // cursor
    J.thr = thr; J.unit_begin = u0; J.unit_end = u1;
    J.out_tgt = W.l_tgt; J.out_oth = W.l_oth; J.out_d = W.l_d; J.out_pos = W.l_pos; J.out_seq = W.l_seq;
    J.tgt_count = W.cnt; J.cursor = W.dev_scalars; J.capacity = (unsigned long long)cap;
    J.overflow = (int *)(W.dev_scalars + 4);
    NND_CUDA(cudaMemsetAsync(W.dev_scalars + 5, 0, sizeof(unsigned long long) * 2, st));
    J.stats = W.dev_scalars + 5;
    NND_STAGE(ctx, NND_ST_JOIN, st);
    NND_CUDA(cudaEventRecord(ctx->ev0, st));
    // schedule options (nnd_ctx): neither changes what is emitted, only how
    const bool rank_scatter = JB && do_apply && ctx->opt_rank_scatter;
    if (JB) {
        JB->thr = thr; JB->batch_begin = u0; JB->batch_end = u1;
        // (records that are exported to their owner process -- do_apply false -- are ranked there: fire-and-forget counting here too)
        JB->out_tgt = W.l_tgt; JB->out_oth = W.l_oth; JB->out_d = W.l_d; JB->out_pos = (rank_scatter || !do_apply) ? nullptr : W.l_pos; JB->out_seq = W.l_seq;
        JB->tgt_count = W.cnt; JB->cursor = W.dev_scalars; JB->capacity = (unsigned long long)cap;
        JB->overflow = (int *)(W.dev_scalars + 4); JB->stats = W.dev_scalars + 5;
        NND_TRY(launch_join_batch(ctx, *JB, st));
    } else {
        NND_TRY(launch_join(ctx, J, u1 - u0, st));
    }
    NND_CUDA(cudaEventRecord(ctx->ev1, st));
    NND_STAGE(ctx, NND_ST_SCATTER, st);
    unsigned long long h[7];
    NND_CUDA(cudaMemcpyAsync(h, W.dev_scalars, sizeof(h), cudaMemcpyDeviceToHost, st));
    NND_CUDA(cudaStreamSynchronize(st));
    if (((int *)&h[4])[0] != 0) {   // the optimistic round did not fit: nothing was applied, the caller splits the round
        NND_CUDA(cudaMemsetAsync(W.dev_scalars + 4, 0, sizeof(unsigned long long), st));
        nnd_set_error("push list overflow (capacity %lld): raise push_capacity", (long long)cap);
        return NND_ERR_OVERFLOW;
    }
    *n_pushed = h[0];
    {
        float ms = 0.f;
        NND_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        double *pf = ctx->prof[J.mode ? 1 : 0];
        pf[0] += ms; pf[1] += 1; pf[2] += (double)h[5]; pf[3] += (double)h[6]; pf[4] += (double)h[0];
        ctx->last_round_rows = h[5]; ctx->last_round_pairs = h[6];
        if (JB && u1 > u0) { double ppb = (double)h[0] / (double)(u1 - u0); if (ppb > ctx->jb_ppb_max) ctx->jb_ppb_max = ppb; }
    }
    if (h[0] == 0 || !do_apply) return NND_OK;
    NND_TRY(nnd_exclusive_scan_u32_padded(ctx, W.cnt, W.off, N, W.scan_tmp, st));  // offsets of power-of-two padded segments
    if (rank_scatter) {   // W.bound is only used by the unbatched plan
        NND_CUDA(cudaMemsetAsync(W.bound, 0, sizeof(uint32_t) * (N + 1), st));
        k_scatter_rank<<<(unsigned)cdiv((int64_t)h[0], 256), 256, 0, st>>>(W.l_tgt, W.l_oth, W.l_d, W.l_seq, h[0], W.off, W.bound, W.rec);
    } else {
        k_scatter<<<(unsigned)cdiv((int64_t)h[0], 256), 256, 0, st>>>(W.l_tgt, W.l_oth, W.l_d, W.l_pos, W.l_seq, h[0], W.off, W.rec);
    }
    NND_LAUNCH(ctx);
    NND_STAGE(ctx, NND_ST_APPLY, st);
    size_t smem = (size_t)APPLY_WARPS * NND_SEG_SMEM * sizeof(Rec16) + (size_t)APPLY_WARPS * (((size_t)P->k * 9 + 15) & ~size_t(15));
    NND_CHECK(smem <= 200 * 1024, NND_ERR_UNSUPPORTED, "apply: k=%d too large", P->k);
    NND_CUDA(cudaFuncSetAttribute(k_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (ctx->opt_carveout) cudaFuncSetAttribute(k_apply, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    const int use_small = 0;
    k_apply<<<(unsigned)cdiv(N, APPLY_WARPS), APPLY_WARPS * 32, smem, st>>>(P->heap_idx, P->heap_dist, P->heap_flag, P->k, N,
                                                                           W.cnt, W.off, W.rec, W.dev_scalars + 1, use_small, (ctx->opt_apply_short ? 1 : 0) | (ctx->heap_trusted ? 2 : 0));
    NND_LAUNCH(ctx);
    NND_STAGE(ctx, NND_ST_OTHER, st);
    NND_CUDA(cudaGetLastError());
    return NND_OK;
}


static size_t join_batch_smem(int metric, int dk, int ld) {
    size_t b = 32 + (size_t)JB_MAXU * 4 * 3 + (size_t)(JB_MAXU + 1) * 4 * 2;
    b += (size_t)JB_SCAP * 4 * (metric == NND_METRIC_ALT_COSINE ? 3 : 2);
    b = (b + 15) & ~size_t(15);
    const int nbuf = (ld <= dk) ? 1 : 2;   // a single chunk needs no second buffer
    b += (size_t)nbuf * JB_SCAP * (dk + 4) * 4 + 16;
    return b;
}
static int launch_join_batch(nnd_ctx *ctx, JoinBatchParams &JB, cudaStream_t st) {
    int64_t nb = JB.batch_end - JB.batch_begin;
    if (nb <= 0) return NND_OK;
    // floats per staged chunk.  16 / 64 / 128 all lost on hardware (1M: profiles/ab_r02_1m_switches.json; 10M: join 513 -> 762 / 732 ms,
    // profiles/ab_r02_10m_call5.json): smaller chunks double the barriers, larger ones halve the resident CTAs
    const int dk = 32;
    size_t smem = join_batch_smem(JB.metric, dk, JB.ld);
    dim3 grid((unsigned)nb);
#define LB(M_, D_)                                                                                                   \
    do {                                                                                                         \
        NND_CUDA(cudaFuncSetAttribute(k_join_batch<M_, D_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_join_batch<M_, D_><<<grid, JB_THREADS, smem, st>>>(JB);                                                     \
    } while (0)
#define LBM(M_) LB(M_, 32)
    if (JB.metric == NND_METRIC_SQEUCLIDEAN) LBM(NND_METRIC_SQEUCLIDEAN);
    else if (JB.metric == NND_METRIC_ALT_COSINE) LBM(NND_METRIC_ALT_COSINE);
    else LBM(NND_METRIC_ALT_DOT);
#undef LBM
#undef LB
    NND_LAUNCH(ctx);
    NND_CUDA(cudaGetLastError());
    return NND_OK;
}

// choose unit blocks so that the worst-case pushes of a block fit in `cap`:
// bound_scan (exclusive prefix of per-unit bounds, copied to host) -> greedy cut points.
static int plan_blocks(nnd_ctx *ctx, const uint32_t *d_bound, int64_t n_units, DescentWs &W, int64_t cap,
                       std::vector<int64_t> &cuts, cudaStream_t st) {
    // coarse: sum bounds per group of G units on the device-side scan, read group prefix on host
    const int64_t G = 1024;
    int64_t ng = cdiv(n_units, G);
    std::vector<uint32_t> hb;  // per-unit bounds are small (<= ~65k); group sums computed on host from the scan
    // exclusive scan may wrap 32 bits for huge inputs; do it per group on the host from raw bounds instead
    hb.resize((size_t)n_units);
    NND_CUDA(cudaMemcpyAsync(hb.data(), d_bound, sizeof(uint32_t) * n_units, cudaMemcpyDeviceToHost, st));
    NND_CUDA(cudaStreamSynchronize(st));
    (void)W; (void)ctx; (void)ng;
    cuts.clear(); cuts.push_back(0);
    int64_t accb = 0;
    for (int64_t g0 = 0; g0 < n_units; g0 += G) {
        int64_t g1 = g0 + G < n_units ? g0 + G : n_units, s = 0;
        for (int64_t i = g0; i < g1; ++i) s += hb[(size_t)i];
        NND_CHECK(s <= cap, NND_ERR_OVERFLOW, "push_capacity %lld too small for %lld units (need %lld)", (long long)cap,
                  (long long)G, (long long)s);
        if (accb + s > cap) { cuts.push_back(g0); accb = 0; }
        accb += s;
    }
    cuts.push_back(n_units);
    return NND_OK;
}

// split leaf init of leaves [l0,l1) (one tree, or a rank's share of one): positions [pos0, pos0 + npos) of leaf_points
static int leaf_split_launch(nnd_ctx *ctx, const nnd_problem *P, const JoinParams &J, int64_t l0, int64_t l1, int64_t pos0, int64_t npos,
                             DescentWs &W, cudaStream_t st) {
    const int64_t nl = l1 - l0;
    if (nl <= 0 || npos <= 0) return NND_OK;
    const size_t lp_smem = (((size_t)P->k * LP_T * 9 + 15) & ~size_t(15)) + (size_t)LP_T * 4;
    NND_CUDA(cudaFuncSetAttribute(k_leaf_push, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lp_smem));
    if (ctx->opt_carveout) cudaFuncSetAttribute(k_leaf_push, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    uint32_t *dsz = W.l_pos, *doff = (uint32_t *)W.l_oth; int32_t *pos2leaf = W.l_tgt; float *D = (float *)W.rec;
    NND_STAGE(ctx, NND_ST_LEAF_DIST, st);
    k_leaf_prep<<<(unsigned)cdiv(nl, 128), 128, 0, st>>>(J.leaf_offsets, l0, l1, pos0, dsz, pos2leaf); NND_LAUNCH(ctx);
    NND_TRY(nnd_exclusive_scan_u32(ctx, dsz, doff, nl, W.scan_tmp, st));
    JoinParams J3 = J; J3.mode = 3; J3.leaf_D = D; J3.leaf_doff = doff; J3.unit_begin = l0; J3.unit_end = l1;
    NND_TRY(launch_join(ctx, J3, nl, st));
    NND_STAGE(ctx, NND_ST_LEAF_PUSH, st);
    k_leaf_push<<<(unsigned)cdiv(npos, LP_T), LP_T, lp_smem, st>>>(J.leaf_points, J.leaf_offsets, l0, pos0, npos, pos2leaf, doff, D, P->heap_idx,
                                                                P->heap_dist, P->heap_flag, P->k);
    NND_LAUNCH(ctx);
    NND_CUDA(cudaGetLastError());
    return NND_OK;
}
static bool leaf_split_fits(const nnd_ctx *ctx, const nnd_problem *P, int mm, int64_t push_capacity) {
    const size_t lp_smem = (((size_t)P->k * LP_T * 9 + 15) & ~size_t(15)) + (size_t)LP_T * 4;
    return ctx->opt_leaf_split && push_capacity >= P->N + 16 && (size_t)push_capacity * 2 * sizeof(Rec16) >= (size_t)P->N * mm * 4 &&
           (double)P->N * mm < 4.0e9 && lp_smem <= 200 * 1024 && join_smem_bytes(mm, P->metric) + 64 <= 200 * 1024;
}

extern "C" int nnd_leaf_init(nnd_ctx *ctx, const nnd_problem *P, const int32_t *leaf_points, const int64_t *leaf_offsets,
                             const int64_t *leaf_offsets_host, int64_t n_leaves, int32_t max_leaf, int64_t push_capacity,
                             void *ws, size_t ws_bytes, int64_t *n_pushes_host, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    NND_CHECK(ctx && P && leaf_points && leaf_offsets && n_leaves >= 0, NND_ERR_INVALID, "nnd_leaf_init: bad args");
    NND_CHECK(max_leaf <= 512, NND_ERR_UNSUPPORTED, "leaf size %d > 512 not supported", max_leaf);
    DescentWs W;
    NND_CHECK(carve_descent_ws(ws, ws_bytes, P->N, P->k, 1, 1, push_capacity, &W, nullptr) == NND_OK, NND_ERR_WORKSPACE,
              "nnd_leaf_init: workspace too small");
    if (n_pushes_host) *n_pushes_host = 0;
    if (n_leaves == 0) return NND_OK;
    NND_CUDA(cudaMemsetAsync(W.dev_scalars, 0, sizeof(unsigned long long) * 16, st));
    // block plan from leaf sizes (host copy of the offsets if given, else read back)
    std::vector<int64_t> offs((size_t)n_leaves + 1);
    if (leaf_offsets_host) memcpy(offs.data(), leaf_offsets_host, sizeof(int64_t) * (n_leaves + 1));
    else {
        NND_CUDA(cudaMemcpyAsync(offs.data(), leaf_offsets, sizeof(int64_t) * (n_leaves + 1), cudaMemcpyDeviceToHost, st));
        NND_CUDA(cudaStreamSynchronize(st));
    }
    // ---- fused path: when the leaves of each tree partition the points (true for every reference
    // forest: offsets hit every multiple of N) a point lies in exactly ONE leaf per tree, so the leaf's
    // own CTA can apply its pushes in place, tree after tree, with no routing at all.
    {
        const int64_t N = P->N, total = offs[n_leaves];
        bool partition = (total % N == 0) && total > 0;
        std::vector<int64_t> tree_cut; tree_cut.push_back(0);
        if (partition) {
            int64_t next = N;
            for (int64_t i = 1; i <= n_leaves && partition; ++i) {
                if (offs[i] == next) { tree_cut.push_back(i); next += N; }
                else if (offs[i] > next) partition = false;
            }
        }
        int mm = max_leaf < 4 ? 4 : max_leaf, al = 0;
        size_t smem = join_smem_bytes(mm, P->metric) + fused_extra_bytes(mm, P->k, &al) + 32;
        const char *force_routed = getenv("NND_LEAF_ROUTED");  // test switch: exercise the routed leaf path
        // NND_LEAF_SPLIT also serves leaves too large for the fused kernel (cfg5: leaf_size 150, k = 30 needs 216 KB fused):
        // the distance-matrix kernel only needs the row buffer
        const bool split_ok = leaf_split_fits(ctx, P, mm, push_capacity);
        if (partition && (smem <= 200 * 1024 || split_ok) && !(force_routed && force_routed[0] == '1')) {
            JoinParams J; memset(&J, 0, sizeof(J));
            J.X = P->X; J.N = N; J.d = P->d; J.ld = P->ld; J.metric = P->metric; J.aux = P->aux;
            J.mode = 2; J.leaf_points = leaf_points; J.leaf_offsets = leaf_offsets; J.m_max = mm; J.slot_stride = mm;
            J.heap_idx = P->heap_idx; J.heap_dist = P->heap_dist; J.heap_flag = P->heap_flag; J.k = P->k; J.d_alias = al;
            // NND_LEAF_SPLIT: distance matrices to global memory (k_join mode 3), pushes by k_leaf_push.  Scratch comes from
            // arrays of the routing workspace that the fused path never touches; fall back to the fused kernel if it is short.
            const bool split = split_ok;
            NND_STAGE(ctx, NND_ST_LEAF_DIST, st);
            NND_CUDA(cudaEventRecord(ctx->ev0, st));
            for (size_t t = 0; t + 1 < tree_cut.size(); ++t) {
                J.unit_begin = tree_cut[t]; J.unit_end = tree_cut[t + 1];
                if (!split) { NND_TRY(launch_join(ctx, J, J.unit_end - J.unit_begin, st)); continue; }
                NND_TRY(leaf_split_launch(ctx, P, J, tree_cut[t], tree_cut[t + 1], offs[tree_cut[t]], offs[tree_cut[t + 1]] - offs[tree_cut[t]], W, st));
            }
            NND_CUDA(cudaGetLastError());
            NND_CUDA(cudaEventRecord(ctx->ev1, st));
            NND_STAGE(ctx, NND_ST_OTHER, st);
            NND_CUDA(cudaStreamSynchronize(st));
            float ms = 0.f;
            NND_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
            double pairs = 0;
            for (int64_t i = 0; i < n_leaves; ++i) { double m = (double)(offs[i + 1] - offs[i]); pairs += m * (m - 1) / 2; }
            double *pf = ctx->prof[1];
            pf[0] += ms; pf[1] += (double)(tree_cut.size() - 1); pf[2] += (double)total; pf[3] += pairs;
            return NND_OK;
        }
    }
    std::vector<int64_t> cuts; cuts.push_back(0);
    int64_t accb = 0;
    for (int64_t i = 0; i < n_leaves; ++i) {
        int64_t m = offs[i + 1] - offs[i], b = m * (m - 1);
        NND_CHECK(m <= max_leaf, NND_ERR_INVALID, "leaf %lld has %lld > max_leaf points", (long long)i, (long long)m);
        NND_CHECK(b <= push_capacity, NND_ERR_OVERFLOW, "push_capacity too small for one leaf");
        if (accb + b > push_capacity) { cuts.push_back(i); accb = 0; }
        accb += b;
    }
    cuts.push_back(n_leaves);
    JoinParams J; memset(&J, 0, sizeof(J));
    J.X = P->X; J.N = P->N; J.d = P->d; J.ld = P->ld; J.metric = P->metric; J.aux = P->aux;
    J.mode = 1; J.leaf_points = leaf_points; J.leaf_offsets = leaf_offsets; J.C = 0;
    J.m_max = max_leaf < 4 ? 4 : max_leaf; J.slot_stride = J.m_max;
    unsigned long long total = 0;
    for (size_t b = 0; b + 1 < cuts.size(); ++b) {
        unsigned long long np = 0;
        if (cuts[b] < cuts[b + 1]) NND_TRY(route_round(ctx, P, J, cuts[b], cuts[b + 1], W, push_capacity, &np, st));
        total += np;
    }
    if (n_pushes_host) *n_pushes_host = (int64_t)total;
    return NND_OK;
}

// ---- join plan: the prepass of one process_candidates call + the list of routing rounds
struct JoinPlan {
    bool batched;
    JoinParams J;
    JoinBatchParams JB;
    std::vector<std::pair<int64_t, int64_t>> rounds;   // unit (or batch) ranges, one per routing round
};

static int build_join_plan(nnd_ctx *ctx, const nnd_problem *P, int32_t C, const int32_t *new_cand, const int32_t *old_cand,
                           int64_t push_capacity, DescentWs &W, cudaStream_t st, JoinPlan &PL) {
    const int64_t N = P->N;
    PL.rounds.clear();
    JoinParams &J = PL.J; memset(&J, 0, sizeof(J));
    J.X = P->X; J.N = N; J.d = P->d; J.ld = P->ld; J.metric = P->metric; J.aux = P->aux;
    J.mode = 0; J.newc = new_cand; J.oldc = old_cand; J.C = C; J.m_max = 2 * C < 4 ? 4 : 2 * C; J.slot_stride = 2 * C;
    J.row_lo = ctx->row_lo; J.row_hi = ctx->row_hi;
    const char *force_old = getenv("NND_JOIN_UNBATCHED");
    // the batched kernel is instantiated for the three f32 dot/difference classes (inner_product shares dot's); everything else
    // runs the generic join
    PL.batched = nnd_metric_class(P->metric) <= NND_METRIC_ALT_DOT && 2 * C <= 60 && !(force_old && force_old[0] == '1');
    NND_STAGE(ctx, NND_ST_JOIN_PREP, st);
    if (PL.batched) {
        // ---- batched path: compact the active rows and their members, cut batches by weight, one CTA per batch
        const int maxphys = ((2 * C + 3) >> 2) << 2;
        const int maxw_t = (jb_tiles(C, 2 * C) * 3 + 1) >> 1;
        const int maxw = maxphys > maxw_t ? maxphys : maxw_t;
        const uint32_t ps = (uint32_t)(JB_SCAP - maxw);                          // weight window per batch
        k_jb_count<<<(unsigned)cdiv(N, 256), 256, 0, st>>>(new_cand, old_cand, C, N, W.jb_flag, W.jb_phys, W.jb_wgt, W.jb_info, ctx->row_lo, ctx->row_hi); NND_LAUNCH(ctx);
        NND_TRY(nnd_exclusive_scan_u32(ctx, W.jb_flag, W.jb_pos, N, W.scan_tmp, st));
        NND_TRY(nnd_exclusive_scan_u32(ctx, W.jb_phys, W.jb_poff, N, W.scan_tmp, st));
        NND_TRY(nnd_exclusive_scan_u32(ctx, W.jb_wgt, W.jb_woff, N, W.scan_tmp, st));
        k_jb_fill<<<(unsigned)cdiv(N, 256), 256, 0, st>>>(new_cand, old_cand, C, N, W.jb_flag, W.jb_pos, W.jb_poff, W.jb_woff, W.jb_info,
                                                          W.act_unit, W.act_info, W.act_poff, W.act_woff, W.members);
        NND_LAUNCH(ctx);
        uint32_t hh[2];
        NND_CUDA(cudaMemcpyAsync(&hh[0], W.jb_pos + N, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        NND_CUDA(cudaMemcpyAsync(&hh[1], W.jb_woff + N, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        NND_CUDA(cudaStreamSynchronize(st));
        const int64_t n_active = hh[0];
        if (n_active > 0) {
            const int64_t n_batches = cdiv((int64_t)hh[1], (int64_t)ps);
            NND_CUDA(cudaMemsetAsync(W.batch_first, 0x7f, sizeof(int32_t) * (n_batches + 2), st));
            k_jb_bounds<<<(unsigned)cdiv(n_active, 256), 256, 0, st>>>(W.act_woff, n_active, ps, W.batch_first); NND_LAUNCH(ctx);
            JoinBatchParams &JB = PL.JB; memset(&JB, 0, sizeof(JB));
            JB.X = P->X; JB.N = N; JB.ld = P->ld; JB.metric = P->metric; JB.aux = P->aux;
            JB.act_unit = W.act_unit; JB.act_info = W.act_info; JB.act_poff = W.act_poff; JB.members = W.members;
            JB.batch_first = W.batch_first; JB.n_active = n_active; JB.n_batches = n_batches; JB.slot_stride = 2 * C;
            // rounds are planned optimistically (16 half-updates per tile slot; the worst case is 32 and the measured
            // average 7); a round that overflows is detected (never silent) and split by the caller
            // ... and from the second iteration of a build on by the measured emission of the previous iteration's rounds
            // (half-updates per batch, x1.3: emission falls from iteration to iteration): fewer, fuller rounds
            int64_t per_batch = (int64_t)JB_TCAP * 16;
            if (ctx->jb_ppb_max > 0.0) {
                int64_t est = (int64_t)(ctx->jb_ppb_max * 1.3) + 64;
                if (est < per_batch) per_batch = est;
            }
            ctx->jb_ppb_max = 0.0;
            NND_CHECK(per_batch <= push_capacity, NND_ERR_OVERFLOW, "push_capacity %lld too small for one batch (%lld)",
                      (long long)push_capacity, (long long)per_batch);
            const int64_t bpb = push_capacity / per_batch;
            for (int64_t b0 = 0; b0 < n_batches; b0 += bpb) PL.rounds.push_back({b0, b0 + bpb < n_batches ? b0 + bpb : n_batches});
        }
    } else {
        k_row_bound<<<(unsigned)cdiv(N, 256), 256, 0, st>>>(new_cand, old_cand, C, N, W.bound, ctx->row_lo, ctx->row_hi); NND_LAUNCH(ctx);
        std::vector<int64_t> cuts;
        NND_TRY(plan_blocks(ctx, W.bound, N, W, push_capacity, cuts, st));
        for (size_t b = 0; b + 1 < cuts.size(); ++b)
            if (cuts[b] < cuts[b + 1]) PL.rounds.push_back({cuts[b], cuts[b + 1]});
    }
    return NND_OK;
}

extern "C" int nnd_process_candidates(nnd_ctx *ctx, const nnd_problem *P, int32_t C, const int32_t *new_cand,
                                      const int32_t *old_cand, int64_t push_capacity, void *ws, size_t ws_bytes,
                                      int64_t *c_host, int64_t *stats_host, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    NND_CHECK(ctx && P && new_cand && old_cand && C > 0 && C <= 128, NND_ERR_INVALID, "nnd_process_candidates: bad args");
    DescentWs W;
    NND_CHECK(carve_descent_ws(ws, ws_bytes, P->N, P->k, C, 1, push_capacity, &W, nullptr) == NND_OK, NND_ERR_WORKSPACE,
              "nnd_process_candidates: workspace too small");
    NND_CUDA(cudaMemsetAsync(W.dev_scalars, 0, sizeof(unsigned long long) * 16, st));
    unsigned long long total = 0, rows_g = 0, pairs_g = 0; int64_t blocks = 0;
    JoinPlan PL;
    NND_TRY(build_join_plan(ctx, P, C, new_cand, old_cand, push_capacity, W, st, PL));
    std::vector<std::pair<int64_t, int64_t>> work(PL.rounds.rbegin(), PL.rounds.rend());   // stack, ascending order on top
    while (!work.empty()) {
        auto r = work.back(); work.pop_back();
        unsigned long long np = 0;
        int rc = route_round(ctx, P, PL.J, r.first, r.second, W, push_capacity, &np, st, PL.batched ? &PL.JB : nullptr);
        if (rc == NND_ERR_OVERFLOW && r.second - r.first > 1) {   // split, keep ascending unit order
            int64_t mid = (r.first + r.second) / 2;
            work.push_back({mid, r.second}); work.push_back({r.first, mid});
            continue;
        }
        if (rc != NND_OK) return rc;
        ++blocks;
        rows_g += ctx->last_round_rows; pairs_g += ctx->last_round_pairs; total += np;
    }
    unsigned long long h[5];
    NND_CUDA(cudaMemcpyAsync(h, W.dev_scalars, sizeof(h), cudaMemcpyDeviceToHost, st));
    NND_CUDA(cudaStreamSynchronize(st));
    if (c_host) *c_host = (int64_t)h[1];
    if (stats_host) { stats_host[0] = (int64_t)rows_g; stats_host[1] = (int64_t)pairs_g; stats_host[2] = (int64_t)total; stats_host[3] = blocks; }
    return NND_OK;
}

// =====================================================================================================
// Sharded build (one process per GPU): the same join, but the half-updates of a round are handed to the
// caller grouped by owner process (row ranges of `owner_block` rows, the reference's own ownership rule
// utils.py:259-264), exchanged with an all-to-all, and applied by the owner.  Every record carries its global
// emission sequence, so the result does not depend on the number of processes.
// =====================================================================================================
struct __align__(8) Rec24 { int32_t tgt, oth; float d; uint32_t pad; unsigned long long seq; };

extern "C" int nnd_join_prepare(nnd_ctx *ctx, const nnd_problem *P, int32_t C, const int32_t *new_cand,
                                const int32_t *old_cand, int64_t push_capacity, void *ws, size_t ws_bytes,
                                int32_t *n_rounds_host, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    NND_CHECK(ctx && P && new_cand && old_cand && n_rounds_host, NND_ERR_INVALID, "nnd_join_prepare: bad args");
    DescentWs W;
    NND_CHECK(carve_descent_ws(ws, ws_bytes, P->N, P->k, C, 1, push_capacity, &W, nullptr) == NND_OK, NND_ERR_WORKSPACE,
              "nnd_join_prepare: workspace too small");
    NND_CUDA(cudaMemsetAsync(W.dev_scalars, 0, sizeof(unsigned long long) * 16, st));
    if (!ctx->plan) { ctx->plan = new JoinPlan(); ctx->plan_free = [](void *p) { delete (JoinPlan *)p; }; }
    JoinPlan &PL = *(JoinPlan *)ctx->plan;
    NND_TRY(build_join_plan(ctx, P, C, new_cand, old_cand, push_capacity, W, st, PL));
    *n_rounds_host = (int32_t)PL.rounds.size();
    return NND_OK;
}

extern "C" int nnd_join_emit_round(nnd_ctx *ctx, const nnd_problem *P, int32_t C, int64_t push_capacity, void *ws,
                                   size_t ws_bytes, int32_t round, int64_t *n_pushed_host, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    NND_CHECK(ctx && P && ctx->plan && n_pushed_host, NND_ERR_INVALID, "nnd_join_emit_round: call nnd_join_prepare first");
    DescentWs W;
    NND_CHECK(carve_descent_ws(ws, ws_bytes, P->N, P->k, C, 1, push_capacity, &W, nullptr) == NND_OK, NND_ERR_WORKSPACE,
              "nnd_join_emit_round: workspace too small");
    JoinPlan &PL = *(JoinPlan *)ctx->plan;
    *n_pushed_host = 0;
    if (round < 0 || round >= (int32_t)PL.rounds.size()) return NND_OK;   // this process has fewer rounds: nothing to emit
    unsigned long long np = 0;
    const std::pair<int64_t, int64_t> r = PL.rounds[round];
    int rc = route_round(ctx, P, PL.J, r.first, r.second, W, push_capacity, &np, st, PL.batched ? &PL.JB : nullptr,
                         /*do_apply=*/false);
    if (rc == NND_ERR_OVERFLOW && r.second - r.first > 1) {
        // the optimistically planned round did not fit the push list: split it in two (the plan grows by one
        // round, ascending unit order kept) and tell the caller to emit the same round index again
        const int64_t mid = (r.first + r.second) / 2;
        PL.rounds[round] = {r.first, mid};
        PL.rounds.insert(PL.rounds.begin() + round + 1, {mid, r.second});
        *n_pushed_host = -1;
        return NND_OK;
    }
    if (rc != NND_OK) return rc;
    *n_pushed_host = (int64_t)np;
    return NND_OK;
}

__global__ void k_owner_count(const int32_t *tgt, unsigned long long n, int64_t owner_block, int R, unsigned long long *counts) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    int o = -1;
    if (i < n) { o = (int)(tgt[i] / owner_block); if (o >= R) o = R - 1; }
    unsigned mk = __match_any_sync(0xffffffffu, o);
    if (o >= 0 && (threadIdx.x & 31) == (__ffs(mk) - 1)) atomicAdd(&counts[o], (unsigned long long)__popc(mk));
}
__global__ void k_owner_scatter(const int32_t *tgt, const int32_t *oth, const float *dd, const unsigned long long *seq,
                                unsigned long long n, int64_t owner_block, int R, unsigned long long *cursors, Rec24 *out) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    int o = -1;
    if (i < n) { o = (int)(tgt[i] / owner_block); if (o >= R) o = R - 1; }
    unsigned mk = __match_any_sync(0xffffffffu, o);
    const int lane = threadIdx.x & 31, leader = __ffs(mk) - 1;
    unsigned long long base = 0;
    if (o >= 0 && lane == leader) base = atomicAdd(&cursors[o], (unsigned long long)__popc(mk));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (o >= 0) {
        Rec24 r; r.tgt = tgt[i]; r.oth = oth[i]; r.d = dd[i]; r.pad = 0; r.seq = seq[i];
        out[base + __popc(mk & ((1u << lane) - 1))] = r;
    }
}

// group the half-updates emitted by the last nnd_join_emit_round by owner process; counts_host[R] receives the
// bucket sizes, `out` (24-byte records, capacity n_pushed) the buckets back to back in rank order
extern "C" int nnd_push_export(nnd_ctx *ctx, const nnd_problem *P, int32_t C, int64_t push_capacity, void *ws, size_t ws_bytes,
                               int64_t n_pushed, int64_t owner_block, int32_t R, void *out, int64_t *counts_host,
                               void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    NND_CHECK(ctx && P && counts_host && R >= 1 && R <= 64 && owner_block > 0, NND_ERR_INVALID, "nnd_push_export: bad args");
    DescentWs W;
    NND_CHECK(carve_descent_ws(ws, ws_bytes, P->N, P->k, C, 1, push_capacity, &W, nullptr) == NND_OK, NND_ERR_WORKSPACE,
              "nnd_push_export: workspace too small");
    for (int r = 0; r < R; ++r) counts_host[r] = 0;
    if (n_pushed <= 0) return NND_OK;
    NND_CHECK(out, NND_ERR_INVALID, "nnd_push_export: out is null");
    unsigned long long *dc = (unsigned long long *)W.rec;   // [128] owner counters + offsets: the routed-record area is free between emit and apply
    NND_CUDA(cudaMemsetAsync(dc, 0, sizeof(unsigned long long) * 128, st));
    k_owner_count<<<(unsigned)cdiv(n_pushed, 256), 256, 0, st>>>(W.l_tgt, (unsigned long long)n_pushed, owner_block, R, dc); NND_LAUNCH(ctx);
    unsigned long long hc[64];
    NND_CUDA(cudaMemcpyAsync(hc, dc, sizeof(unsigned long long) * R, cudaMemcpyDeviceToHost, st));
    NND_CUDA(cudaStreamSynchronize(st));
    unsigned long long offs[64], acc = 0;
    for (int r = 0; r < R; ++r) { offs[r] = acc; acc += hc[r]; counts_host[r] = (int64_t)hc[r]; }
    NND_CUDA(cudaMemcpyAsync(dc + 64, offs, sizeof(unsigned long long) * R, cudaMemcpyHostToDevice, st));
    k_owner_scatter<<<(unsigned)cdiv(n_pushed, 256), 256, 0, st>>>(W.l_tgt, W.l_oth, W.l_d, W.l_seq, (unsigned long long)n_pushed,
                                                                  owner_block, R, dc + 64, (Rec24 *)out);
    NND_LAUNCH(ctx);
    NND_CUDA(cudaStreamSynchronize(st));   // offs[] is a stack array
    NND_CUDA(cudaGetLastError());
    return NND_OK;
}

// cnt / off are indexed by (target - row_lo): only the rows this process owns are histogrammed and replayed
__global__ void k_rec_count(const Rec24 *in, unsigned long long n, uint32_t *cnt, uint32_t *pos, int64_t row_lo, int64_t row_hi,
                            unsigned long long *bad) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t t = in[i].tgt;
    if (t < row_lo || t >= row_hi) { pos[i] = 0xffffffffu; atomicAdd(bad, 1ULL); return; }
    pos[i] = atomicAdd(&cnt[t - row_lo], 1u);
}
__global__ void k_rec_scatter(const Rec24 *in, const uint32_t *pos, unsigned long long n, const uint32_t *off, Rec16 *rec,
                              int64_t row_lo) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (pos[i] == 0xffffffffu) return;
    Rec24 a = in[i];
    Rec16 r; r.key = a.seq; r.oth = a.oth; r.d = a.d;
    rec[(size_t)off[a.tgt - row_lo] + pos[i]] = r;
}

// apply received half-updates (any order) to the local heap rows: counting sort by target + ordered replay.
// c_host = accepted pushes (utils.py:661-733 return value, local part)
extern "C" int nnd_apply_records(nnd_ctx *ctx, const nnd_problem *P, int32_t C, const void *records, int64_t n,
                                 int64_t push_capacity, void *ws, size_t ws_bytes, int64_t *c_host, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    NND_CHECK(ctx && P && c_host && n >= 0, NND_ERR_INVALID, "nnd_apply_records: bad args");
    NND_CHECK(n <= push_capacity, NND_ERR_OVERFLOW, "nnd_apply_records: %lld records exceed push_capacity %lld", (long long)n,
              (long long)push_capacity);
    DescentWs W;
    NND_CHECK(carve_descent_ws(ws, ws_bytes, P->N, P->k, C, 1, push_capacity, &W, nullptr) == NND_OK, NND_ERR_WORKSPACE,
              "nnd_apply_records: workspace too small");
    *c_host = 0;
    if (n == 0) return NND_OK;
    // only the rows this process owns ([row_lo,row_hi), nnd_set_row_range) can be targets: histogram, scan and
    // replay over that range instead of all N rows
    const int64_t N = P->N;
    const int64_t lo = ctx->row_lo > 0 ? (ctx->row_lo < N ? ctx->row_lo : N) : 0;
    const int64_t hi = ctx->row_hi < N ? (ctx->row_hi > lo ? ctx->row_hi : lo) : N;
    const int64_t Nr = hi - lo;
    NND_CHECK(Nr > 0, NND_ERR_INVALID, "nnd_apply_records: %lld records but the row range [%lld,%lld) is empty", (long long)n,
              (long long)lo, (long long)hi);
    NND_CUDA(cudaMemsetAsync(W.cnt, 0, sizeof(uint32_t) * (Nr + 1), st));
    NND_CUDA(cudaMemsetAsync(W.dev_scalars + 1, 0, sizeof(unsigned long long), st));
    NND_CUDA(cudaMemsetAsync(W.dev_scalars + 9, 0, sizeof(unsigned long long), st));
    k_rec_count<<<(unsigned)cdiv(n, 256), 256, 0, st>>>((const Rec24 *)records, (unsigned long long)n, W.cnt, W.l_pos, lo, hi, W.dev_scalars + 9); NND_LAUNCH(ctx);
    NND_TRY(nnd_exclusive_scan_u32_padded(ctx, W.cnt, W.off, Nr, W.scan_tmp, st));
    k_rec_scatter<<<(unsigned)cdiv(n, 256), 256, 0, st>>>((const Rec24 *)records, W.l_pos, (unsigned long long)n, W.off, W.rec, lo); NND_LAUNCH(ctx);
    size_t smem = (size_t)APPLY_WARPS * NND_SEG_SMEM * sizeof(Rec16) + (size_t)APPLY_WARPS * (((size_t)P->k * 9 + 15) & ~size_t(15));
    NND_CUDA(cudaFuncSetAttribute(k_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int skip_small = 0;
    k_apply<<<(unsigned)cdiv(Nr, APPLY_WARPS), APPLY_WARPS * 32, smem, st>>>(P->heap_idx + lo * P->k, P->heap_dist + lo * P->k,
                                                                            P->heap_flag + lo * P->k, P->k, Nr, W.cnt, W.off, W.rec,
                                                                            W.dev_scalars + 1, skip_small, (ctx->opt_apply_short ? 1 : 0) | (ctx->heap_trusted ? 2 : 0));
    NND_LAUNCH(ctx);
    unsigned long long hc = 0, hbad = 0;
    NND_CUDA(cudaMemcpyAsync(&hc, W.dev_scalars + 1, sizeof(hc), cudaMemcpyDeviceToHost, st));
    NND_CUDA(cudaMemcpyAsync(&hbad, W.dev_scalars + 9, sizeof(hbad), cudaMemcpyDeviceToHost, st));
    NND_CUDA(cudaStreamSynchronize(st));
    *c_host = (int64_t)hc;
    NND_CUDA(cudaGetLastError());
    NND_CHECK(hbad == 0, NND_ERR_INVALID, "nnd_apply_records: %llu records target rows outside this process's range [%lld,%lld)",
              hbad, (long long)lo, (long long)hi);
    return NND_OK;
}

// fused leaf join + apply on leaves [l0,l1) of ONE tree (caller guarantees the leaves are disjoint point sets)
extern "C" int nnd_leaf_fused_range(nnd_ctx *ctx, const nnd_problem *P, const int32_t *leaf_points, const int64_t *leaf_offsets,
                                    int64_t l0, int64_t l1, int32_t max_leaf, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    NND_CHECK(ctx && P && leaf_points && leaf_offsets && l0 >= 0 && l1 >= l0, NND_ERR_INVALID, "nnd_leaf_fused_range: bad args");
    if (l1 == l0) return NND_OK;
    int mm = max_leaf < 4 ? 4 : max_leaf, al = 0;
    size_t smem = join_smem_bytes(mm, P->metric) + fused_extra_bytes(mm, P->k, &al) + 32;
    NND_CHECK(smem <= 200 * 1024, NND_ERR_UNSUPPORTED, "fused leaf join needs %zu B of shared memory (leaf %d, k %d)", smem, mm, P->k);
    JoinParams J; memset(&J, 0, sizeof(J));
    J.X = P->X; J.N = P->N; J.d = P->d; J.ld = P->ld; J.metric = P->metric; J.aux = P->aux;
    J.mode = 2; J.leaf_points = leaf_points; J.leaf_offsets = leaf_offsets; J.m_max = mm; J.slot_stride = mm;
    J.heap_idx = P->heap_idx; J.heap_dist = P->heap_dist; J.heap_flag = P->heap_flag; J.k = P->k; J.d_alias = al;
    J.unit_begin = l0; J.unit_end = l1;
    return launch_join(ctx, J, l1 - l0, st);
}

// the same leaves through the split leaf init (NND_LEAF_SPLIT; k_join mode 3 + k_leaf_push) when it is enabled and its scratch
// fits the routing workspace, else through the fused kernel.  pos0 / npos = the position range of the leaves in leaf_points
// (the caller knows the offsets on the host).
extern "C" int nnd_leaf_range(nnd_ctx *ctx, const nnd_problem *P, const int32_t *leaf_points, const int64_t *leaf_offsets, int64_t l0,
                              int64_t l1, int64_t pos0, int64_t npos, int32_t max_leaf, int64_t push_capacity, void *ws, size_t ws_bytes,
                              void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    NND_CHECK(ctx && P && leaf_points && leaf_offsets && l0 >= 0 && l1 >= l0 && pos0 >= 0 && npos >= 0, NND_ERR_INVALID, "nnd_leaf_range: bad args");
    const int mm = max_leaf < 4 ? 4 : max_leaf;
    DescentWs W;
    if (!ws || carve_descent_ws(ws, ws_bytes, P->N, P->k, 1, 1, push_capacity, &W, nullptr) != NND_OK || !leaf_split_fits(ctx, P, mm, push_capacity))
        return nnd_leaf_fused_range(ctx, P, leaf_points, leaf_offsets, l0, l1, max_leaf, stream);
    JoinParams J; memset(&J, 0, sizeof(J));
    J.X = P->X; J.N = P->N; J.d = P->d; J.ld = P->ld; J.metric = P->metric; J.aux = P->aux;
    J.leaf_points = leaf_points; J.leaf_offsets = leaf_offsets; J.m_max = mm; J.slot_stride = mm;
    J.heap_idx = P->heap_idx; J.heap_dist = P->heap_dist; J.heap_flag = P->heap_flag; J.k = P->k;
    return leaf_split_launch(ctx, P, J, l0, l1, pos0, npos, W, st);
}

// pack / unpack heap rows listed in `rows` (sharded build: exchange of the rows a process has updated)
__global__ void k_heap_pack(const int32_t *hi, const float *hd, const uint8_t *hf, int k, const int32_t *rows, int64_t n,
                            int32_t *oi, float *od, uint8_t *of, int unpack, int32_t *whi, float *whd, uint8_t *whf) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * k) return;
    int64_t r = e / k; int j = (int)(e - r * k);
    int64_t row = rows ? (int64_t)rows[r] : r;
    if (!unpack) { oi[e] = hi[row * k + j]; od[e] = hd[row * k + j]; of[e] = hf[row * k + j]; }
    else { whi[row * k + j] = oi[e]; whd[row * k + j] = od[e]; whf[row * k + j] = of[e]; }
}
extern "C" int nnd_heap_pack_rows(nnd_ctx *ctx, const nnd_problem *P, const int32_t *rows, int64_t n, int32_t *out_idx,
                                  float *out_dist, uint8_t *out_flag, int32_t unpack, void *stream) {
    NND_CHECK(ctx && P && out_idx && out_dist && out_flag && n >= 0, NND_ERR_INVALID, "nnd_heap_pack_rows: bad args");
    if (n == 0) return NND_OK;
    k_heap_pack<<<(unsigned)cdiv(n * P->k, 256), 256, 0, (cudaStream_t)stream>>>(P->heap_idx, P->heap_dist, P->heap_flag, P->k, rows, n,
                                                                              out_idx, out_dist, out_flag, unpack, P->heap_idx,
                                                                              P->heap_dist, P->heap_flag);
    NND_LAUNCH(ctx);
    NND_CUDA(cudaGetLastError());
    return NND_OK;
}

// pynndescent_.py:266-320 nn_descent_internal
extern "C" int nnd_descent(nnd_ctx *ctx, const nnd_problem *P, int32_t C, const int64_t *rng_state_host, int32_t T,
                           int32_t n_iters, double delta, int64_t push_capacity, void *ws, size_t ws_bytes,
                           int32_t *iters_run_host, int64_t *c_hist_host, int64_t *stats_host, void *stream) {
    NND_CHECK(ctx && P && rng_state_host && ws, NND_ERR_INVALID, "nnd_descent: bad args");
    DescentWs W;
    NND_CHECK(carve_descent_ws(ws, ws_bytes, P->N, P->k, C, T, push_capacity, &W, nullptr) == NND_OK, NND_ERR_WORKSPACE,
              "nnd_descent: workspace too small");
    int it = 0;
    int64_t tot[4] = {0, 0, 0, 0};
    for (it = 0; it < n_iters; ++it) {
        NND_TRY(nnd_build_candidates(ctx, P, C, rng_state_host, T, it > 0, W.newc, W.oldc, W.cand_base, W.cand_bytes, stream));
        int64_t c = 0, stt[4];
        // process_candidates re-carves the same workspace: identical layout, so the candidate
        // arrays and rng tables carved above stay where they are
        NND_TRY(nnd_process_candidates(ctx, P, C, W.newc, W.oldc, push_capacity, ws, ws_bytes, &c, stt, stream));
        for (int q = 0; q < 4; ++q) tot[q] += stt[q];
        if (c_hist_host) c_hist_host[it] = c;
        if ((double)c <= delta * (double)P->k * (double)P->N) { ++it; break; }
    }
    if (iters_run_host) *iters_run_host = it;
    if (stats_host) for (int q = 0; q < 4; ++q) stats_host[q] = tot[q];
    return NND_OK;
}
