// k_search_graph.cu -- the sparse-matrix part of _init_search_graph (pynndescent_.py:1521-1651) on the device.
//
// Reference sequence (scipy on the host): forward-diversified graph -> CSR, second diversification pass, G = max(G, G^T)
// (:1601), setdiag(0) (:1605), degree_prune (:728-760: in a row with more than max_degree entries everything larger than
// the (max_degree+1)-th smallest weight is dropped, ties with it stay), != 0 (:1613) and the permutation of rows and
// columns to RP-tree leaf order (:1629-1634).  Every step is row-local once the reverse edges of a row are known, so:
//
//   k_sg_indeg / scan / k_sg_fill   counting sort of the edges by target  (reverse adjacency, order arbitrary)
//   k_sg_sizes / scan               per-row scratch segment = forward + reverse entries (power of two beyond SG_CAP)
//   k_sg_rows   (warp per row)      sort by neighbour id, keep the larger weight of a pair of duplicates (= maximum),
//                                   drop the diagonal, degree-prune through a sort by weight, rename by leaf order, sort by
//                                   the new id -> the finished row, left in the scratch segment
//   scan / k_sg_write               CSR assembly in leaf order
//   k_sg_min                        min_distance (:1539) over the FIRST-pass graph
//
// pynndescent_b200/prepare.py:search_graph_bookkeeping is the host restatement of the same sequence (pinned to the
// reference's structures in tests/test_prepare_host.py); tests/test_kernel_emulation.py runs these kernels against it.
//
// STATUS: written after round 1's GPU budget was spent: compiled and emulated, not yet run on hardware, hence OFF by
// default (NND_PREPARE_GPU=1 makes prepare() use it).
#include "nnd_common.cuh"

#define SG_WARPS 4
#define SG_CAP 256   // entries a warp sorts in shared memory; larger rows are sorted in their global scratch segment

__host__ __device__ __forceinline__ uint32_t sg_padded(uint32_t m) { return m <= SG_CAP ? m : next_pow2_u32(m); }

__global__ void k_sg_min(const int32_t *di, const float *dd, int64_t n, unsigned *min_bits) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n && di[e] >= 0) atomicMin(min_bits, __float_as_uint(dd[e]));   // weights are positive: float order = bit order
}
__global__ void k_sg_inv(const int32_t *vertex_order, int64_t N, int32_t *inv) {
    int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a < N) inv[vertex_order[a]] = (int32_t)a;
}
__global__ void k_sg_indeg(const int32_t *di2, int k, int64_t N, uint32_t *indeg) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N * k) return;
    const int32_t j = di2[e];
    if (j >= 0 && j != (int32_t)(e / k)) atomicAdd(&indeg[j], 1u);
}
__global__ void k_sg_fill(const int32_t *di2, const float *dd, int k, int64_t N, const uint32_t *roff, uint32_t *cursor,
                          unsigned long long *rev) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N * k) return;
    const int32_t j = di2[e], i = (int32_t)(e / k);
    if (j < 0 || j == i) return;
    const uint32_t slot = atomicAdd(&cursor[j], 1u);
    rev[(size_t)roff[j] + slot] = ((unsigned long long)(uint32_t)i << 32) | (unsigned long long)__float_as_uint(dd[e]);
}
__global__ void k_sg_sizes(const int32_t *di2, int k, int64_t N, const uint32_t *indeg, uint32_t *m_out, uint32_t *pad_out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    uint32_t m = indeg[i];
    for (int c = 0; c < k; ++c) { const int32_t j = di2[i * k + c]; m += (j >= 0 && j != (int32_t)i); }
    m_out[i] = m; pad_out[i] = sg_padded(m);
}

// order-preserving compaction of a[0..n) by `keep` (evaluated on the element), in place, by one warp
template <class KEEP>
__device__ __forceinline__ int sg_compact(unsigned long long *a, int n, int lane, KEEP keep) {
    int out = 0;
    for (int b0 = 0; b0 < n; b0 += 32) {
        const int t = b0 + lane;
        unsigned long long v = 0; bool kp = false;
        if (t < n) { v = a[t]; kp = keep(v, t); }
        const unsigned mk = __ballot_sync(0xffffffffu, kp);
        __syncwarp();                                     // everyone has read its element before anything is overwritten
        if (kp) a[out + __popc(mk & ((1u << lane) - 1))] = v;
        out += __popc(mk);
        __syncwarp();
    }
    return out;
}
__device__ __forceinline__ void sg_sort(unsigned long long *a, int n, int lane) {   // pads to a power of two with ~0
    const int p2 = (int)next_pow2_u32((uint32_t)(n > 1 ? n : 1));
    for (int t = n + lane; t < p2; t += 32) a[t] = ~0ULL;
    __syncwarp();
    if (p2 > 1) warp_bitonic_sort_u64(a, p2, lane);
    __syncwarp();
}

__global__ void __launch_bounds__(SG_WARPS * 32) k_sg_rows(const int32_t *di2, const float *dd, int k, int64_t N,
                                                          const uint32_t *indeg, const uint32_t *roff,
                                                          const unsigned long long *rev, const uint32_t *m_arr,
                                                          const uint32_t *soff, unsigned long long *scratch, const int32_t *inv,
                                                          int max_degree, uint32_t *cnt_new) {
    __shared__ unsigned long long sbuf[SG_WARPS][SG_CAP];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * SG_WARPS + warp;
    if (i >= N) return;
    const int m = (int)m_arr[i];
    unsigned long long *seg = scratch + soff[i];
    if (m == 0) { if (lane == 0) cnt_new[inv[i]] = 0; return; }
    unsigned long long *a = m <= SG_CAP ? sbuf[warp] : seg;
    // ---- gather: forward entries (id, w) then reverse entries; key = id << 32 | float bits (weights > 0)
    int nf = 0;
    for (int c0 = 0; c0 < k; c0 += 32) {
        const int c = c0 + lane;
        int32_t j = -1; float w = 0.f;
        if (c < k) { j = di2[i * k + c]; w = dd[i * k + c]; }
        const bool ok = j >= 0 && j != (int32_t)i;
        const unsigned mk = __ballot_sync(0xffffffffu, ok);
        if (ok) a[nf + __popc(mk & ((1u << lane) - 1))] = ((unsigned long long)(uint32_t)j << 32) | (unsigned long long)__float_as_uint(w);
        nf += __popc(mk);
    }
    const int nr = (int)indeg[i];
    for (int t = lane; t < nr; t += 32) a[nf + t] = rev[(size_t)roff[i] + t];
    __syncwarp();
    // ---- maximum(G, G^T): sort by (id, w); of equal ids keep the last (= the larger weight)
    sg_sort(a, m, lane);
    int u = sg_compact(a, m, lane, [&](unsigned long long v, int t) { return t == m - 1 || (uint32_t)(a[t + 1] >> 32) != (uint32_t)(v >> 32); });
    __syncwarp();
    // ---- degree prune (:728-760)
    if (u > max_degree) {
        for (int t = lane; t < u; t += 32) { const unsigned long long v = a[t]; a[t] = (v << 32) | (v >> 32); }   // (w, id)
        __syncwarp();
        sg_sort(a, u, lane);
        const uint32_t cut = (uint32_t)(a[max_degree] >> 32);
        __syncwarp();
        u = sg_compact(a, u, lane, [&](unsigned long long v, int) { return (uint32_t)(v >> 32) <= cut; });
        __syncwarp();
        for (int t = lane; t < u; t += 32) a[t] = (unsigned long long)(uint32_t)inv[(uint32_t)(a[t] & 0xFFFFFFFFULL)];
    } else {
        for (int t = lane; t < u; t += 32) a[t] = (unsigned long long)(uint32_t)inv[(uint32_t)(a[t] >> 32)];
    }
    __syncwarp();
    // ---- leaf order (:1629-1634): renamed ids ascending; the finished row stays in the scratch segment as int32
    sg_sort(a, u, lane);
    int32_t *out = (int32_t *)seg;
    for (int b0 = 0; b0 < u; b0 += 32) {                  // u64 -> i32 in place: element t is written below where it was read
        const int t = b0 + lane;
        int32_t v = 0;
        if (t < u) v = (int32_t)(uint32_t)a[t];
        __syncwarp();
        if (t < u) out[t] = v;
        __syncwarp();
    }
    if (lane == 0) cnt_new[inv[i]] = (uint32_t)u;
}

__global__ void __launch_bounds__(SG_WARPS * 32) k_sg_write(const int32_t *vertex_order, int64_t N, const uint32_t *soff,
                                                           const unsigned long long *scratch, const uint32_t *cnt_new,
                                                           const uint32_t *indptr, int32_t *indptr_out, int32_t *indices) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t a = (int64_t)blockIdx.x * SG_WARPS + warp;
    if (a > N) return;
    if (lane == 0) indptr_out[a] = (int32_t)indptr[a];
    if (a == N) return;
    const int32_t *src = (const int32_t *)(scratch + soff[vertex_order[a]]);
    const uint32_t n = cnt_new[a], o = indptr[a];
    for (uint32_t t = lane; t < n; t += 32) indices[o + t] = src[t];
}

// ------------------------------------------------------------------------------------------------ host
struct SgWs { uint32_t *indeg, *roff, *cursor, *m, *pad, *soff, *cnt_new, *indptr, *scan_tmp; int32_t *inv; unsigned *scalars;
              unsigned long long *rev, *scratch; };
static int carve_sg(void *ws, size_t ws_bytes, int64_t N, int k, SgWs *W, size_t *need) {
    WsAlloc A(ws, ws_bytes);
    const int64_t Nb = N > 16 ? N : 16;
    W->indeg = A.get<uint32_t>(Nb + 1); W->roff = A.get<uint32_t>(Nb + 2); W->cursor = A.get<uint32_t>(Nb + 1);
    W->m = A.get<uint32_t>(Nb + 1); W->pad = A.get<uint32_t>(Nb + 1); W->soff = A.get<uint32_t>(Nb + 2);
    W->cnt_new = A.get<uint32_t>(Nb + 1); W->indptr = A.get<uint32_t>(Nb + 2); W->inv = A.get<int32_t>(Nb + 1);
    W->scan_tmp = A.get<uint32_t>(nnd_scan_tmp_elems(Nb + 2));
    W->scalars = A.get<unsigned>(16);
    W->rev = A.get<unsigned long long>((size_t)Nb * k + 16);
    W->scratch = A.get<unsigned long long>((size_t)4 * Nb * k + SG_CAP);   // sum of padded (forward + reverse) row sizes < 2 * 2Nk
    if (need) *need = A.off;
    return A.ok ? NND_OK : NND_ERR_WORKSPACE;
}
extern "C" int nnd_search_graph_workspace_bytes(int64_t N, int32_t k, size_t *bytes) {
    NND_CHECK(bytes && N > 0 && k > 0, NND_ERR_INVALID, "nnd_search_graph_workspace_bytes: bad args");
    SgWs W; size_t need = 0;
    carve_sg(nullptr, 0, N, k, &W, &need);
    *bytes = need + 4096;
    return NND_OK;
}
extern "C" int nnd_search_graph_build(nnd_ctx *ctx, const int32_t *di, const float *dd, const int32_t *di2, const float *dd2, int64_t N, int32_t k,
                                      const int32_t *vertex_order, int32_t max_degree, int32_t *indptr_out, int32_t *indices_out,
                                      int64_t capacity, int64_t *nnz_host, float *min_distance_host, void *ws, size_t ws_bytes,
                                      void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    NND_CHECK(ctx && di && di2 && dd && dd2 && vertex_order && indptr_out && indices_out && nnz_host && min_distance_host && N > 0 && k > 0 &&
                  max_degree >= 0, NND_ERR_INVALID, "nnd_search_graph_build: bad args");
    NND_CHECK((double)N * k * 4 < 4.0e9, NND_ERR_UNSUPPORTED, "nnd_search_graph_build: N*k too large for 32-bit offsets");
    SgWs W;
    NND_CHECK(carve_sg(ws, ws_bytes, N, k, &W, nullptr) == NND_OK, NND_ERR_WORKSPACE, "nnd_search_graph_build: workspace too small");
    const unsigned ge = (unsigned)cdiv(N * k, 256), gn = (unsigned)cdiv(N, 256);
    NND_CUDA(cudaMemsetAsync(W.indeg, 0, sizeof(uint32_t) * (N + 1), st));
    NND_CUDA(cudaMemsetAsync(W.cursor, 0, sizeof(uint32_t) * (N + 1), st));
    NND_CUDA(cudaMemsetAsync(W.scalars, 0xff, sizeof(unsigned), st));
    k_sg_min<<<ge, 256, 0, st>>>(di, dd, N * k, W.scalars); NND_LAUNCH(ctx);
    k_sg_inv<<<gn, 256, 0, st>>>(vertex_order, N, W.inv); NND_LAUNCH(ctx);
    k_sg_indeg<<<ge, 256, 0, st>>>(di2, k, N, W.indeg); NND_LAUNCH(ctx);
    NND_TRY(nnd_exclusive_scan_u32(ctx, W.indeg, W.roff, N, W.scan_tmp, st));
    k_sg_fill<<<ge, 256, 0, st>>>(di2, dd2, k, N, W.roff, W.cursor, W.rev); NND_LAUNCH(ctx);
    k_sg_sizes<<<gn, 256, 0, st>>>(di2, k, N, W.indeg, W.m, W.pad); NND_LAUNCH(ctx);
    NND_TRY(nnd_exclusive_scan_u32(ctx, W.pad, W.soff, N, W.scan_tmp, st));
    k_sg_rows<<<(unsigned)cdiv(N, SG_WARPS), SG_WARPS * 32, 0, st>>>(di2, dd2, k, N, W.indeg, W.roff, W.rev, W.m, W.soff, W.scratch, W.inv,
                                                                    max_degree, W.cnt_new);
    NND_LAUNCH(ctx);
    NND_TRY(nnd_exclusive_scan_u32(ctx, W.cnt_new, W.indptr, N, W.scan_tmp, st));
    uint32_t nnz = 0; unsigned mb = 0;
    NND_CUDA(cudaMemcpyAsync(&nnz, W.indptr + N, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    NND_CUDA(cudaMemcpyAsync(&mb, W.scalars, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    NND_CUDA(cudaStreamSynchronize(st));
    *nnz_host = (int64_t)nnz;
    memcpy(min_distance_host, &mb, sizeof(float));
    NND_CHECK((int64_t)nnz <= capacity, NND_ERR_OVERFLOW, "nnd_search_graph_build: %u entries exceed the output capacity %lld", nnz,
              (long long)capacity);
    k_sg_write<<<(unsigned)cdiv(N + 1, SG_WARPS), SG_WARPS * 32, 0, st>>>(vertex_order, N, W.soff, W.scratch, W.cnt_new, W.indptr, indptr_out,
                                                                        indices_out);
    NND_LAUNCH(ctx);
    NND_CUDA(cudaGetLastError());
    return NND_OK;
}
