// k_prepare.cu -- GPU parts of search-graph preparation (pynndescent_.py:1333-1662 _init_search_graph):
//   * the graph-informed HUB TREE (rp_trees.py:714-1309 make_hub_tree: the three highest in-degree members of a
//     node define three candidate hyperplanes, the most balanced one wins, a node whose best balance is < 0.1
//     stays a leaf).  It uses no randomness (the rng fallback always ends in a leaf), so the level-synchronous
//     build below produces THE SAME tree as the reference's recursion;
//   * forward diversification (pynndescent_.py:369-403 diversify with prune_probability = 1) and the row-wise
//     diversify_csr (:549-588), both "drop neighbour j if a kept neighbour c is closer to j than the row's point".
// The sparse-matrix bookkeeping in between (symmetrise, degree prune, re-order) is done by the host with scipy,
// exactly the calls the reference makes.
#include <vector>

#include "nnd_common.cuh"

// ------------------------------------------------------------------ canonical warp reductions (as in k_forest.cu)
__device__ __forceinline__ float hub_lane_dot4(const float4 &h, const float *__restrict__ x, int ld, int lane, const float *hfull) {
    float acc = 0.0f;
    if (ld <= 128) {
        if (lane * 4 < ld) { float4 b = *(const float4 *)(x + lane * 4);
            acc = __fmaf_rn(h.x, b.x, acc); acc = __fmaf_rn(h.y, b.y, acc); acc = __fmaf_rn(h.z, b.z, acc); acc = __fmaf_rn(h.w, b.w, acc); }
    } else {
        for (int c = lane * 4; c < ld; c += 128) {
            float4 a = *(const float4 *)(hfull + c), b = *(const float4 *)(x + c);
            acc = __fmaf_rn(a.x, b.x, acc); acc = __fmaf_rn(a.y, b.y, acc); acc = __fmaf_rn(a.z, b.z, acc); acc = __fmaf_rn(a.w, b.w, acc);
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
    return acc;
}
__device__ __forceinline__ float hub_lane_norm(const float *__restrict__ x, int ld, int lane) {
    double acc = 0.0;
    for (int c = lane * 4; c < ld; c += 128) {
        float4 a = *(const float4 *)(x + c);
        acc = __dadd_rn(acc, (double)__fmul_rn(a.x, a.x)); acc = __dadd_rn(acc, (double)__fmul_rn(a.y, a.y));
        acc = __dadd_rn(acc, (double)__fmul_rn(a.z, a.z)); acc = __dadd_rn(acc, (double)__fmul_rn(a.w, a.w));
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
    return (float)__dsqrt_rn(acc);
}

__global__ void k_hub_degrees(const int32_t *ng, int64_t N, int k, int32_t *deg) {   // rp_trees.py:714-744
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N * k) return;
    int32_t v = ng[e];
    if (v >= 0 && v < N) atomicAdd(&deg[v], 1);
}
__global__ void k_hub_init(int32_t *perm, int32_t *cand, int64_t N, int active) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) { perm[i] = (int32_t)i; cand[i] = active ? 0 : -1; }
}

// one warp per candidate node: top-3 hubs (rp_trees.py:747-803: degree descending, earlier member first on ties)
// and the three candidate hyperplanes (pairs (0,1),(0,2),(1,2); :842-853 / :969-990)
__global__ void k_hub_top3(const float *X, int ld, int angular, int M, const int32_t *cstart, const int32_t *csize,
                           const int32_t *perm, const int32_t *deg, float *hp3, float *off3) {
    extern __shared__ float s_t[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + warp;
    if (c >= M) return;
    const int start = cstart[c], n = csize[c];
    int top_pos[3] = {-1, -1, -1};
    for (int r = 0; r < 3; ++r) {
        long long best = -1;   // (deg << 32) | (0x7fffffff - rel position): larger is better
        for (int i = lane; i < n; i += 32) {
            if (i == top_pos[0] || i == top_pos[1]) continue;
            long long key = ((long long)deg[perm[start + i]] << 32) | (long long)(0x7fffffff - i);
            if (key > best) best = key;
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) { long long other = __shfl_xor_sync(0xffffffffu, best, o); if (other > best) best = other; }
        top_pos[r] = best < 0 ? -1 : (int)(0x7fffffff - (int)(best & 0xffffffffLL));
    }
    float *t = s_t + (size_t)warp * ld;
    int pi = 0;
    for (int hi = 0; hi < 3; ++hi)
        for (int hj = hi + 1; hj < 3; ++hj, ++pi) {
            float *h = hp3 + ((size_t)c * 3 + pi) * ld;
            if (top_pos[hi] < 0 || top_pos[hj] < 0) { for (int e = lane; e < ld; e += 32) h[e] = 0.f; if (lane == 0) off3[c * 3 + pi] = CUDART_NAN_F; continue; }
            const float *xl = X + (int64_t)perm[start + top_pos[hi]] * ld, *xr = X + (int64_t)perm[start + top_pos[hj]] * ld;
            if (!angular) {
                for (int e = lane; e < ld; e += 32) { float he = __fsub_rn(xl[e], xr[e]); h[e] = he; t[e] = __fmul_rn(he, __fadd_rn(xl[e], xr[e])); }
                __syncwarp();
                if (lane == 0) { float ofs = 0.0f; for (int e = 0; e < ld; ++e) ofs = __fsub_rn(ofs, __fmul_rn(0.5f, t[e])); off3[c * 3 + pi] = ofs; }
                __syncwarp();
            } else {
                float ln = hub_lane_norm(xl, ld, lane), rn = hub_lane_norm(xr, ld, lane);
                if (fabs((double)ln) < 1e-8) ln = 1.0f;
                if (fabs((double)rn) < 1e-8) rn = 1.0f;
                for (int e = lane; e < ld; e += 32) t[e] = __fsub_rn(__fdiv_rn(xl[e], ln), __fdiv_rn(xr[e], rn));
                __syncwarp();
                float hn = hub_lane_norm(t, ld, lane);
                if (fabs((double)hn) < 1e-8) hn = 1.0f;
                for (int e = lane; e < ld; e += 32) h[e] = __fdiv_rn(t[e], hn);
                if (lane == 0) off3[c * 3 + pi] = 0.0f;
                __syncwarp();
            }
        }
}

// side of every member for the three candidate hyperplanes (:857-876): >EPS left, <-EPS right, else i % 2
#define HUB_RUN 16
__global__ void __launch_bounds__(256) k_hub_side(const float *X, int ld, int64_t N, const int32_t *perm, const int32_t *cand,
                                                  const int32_t *cstart, const float *hp3, const float *off3, uint8_t *side3) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t i0 = w * HUB_RUN;
    if (i0 >= N) return;
    const int cnt = (int)min((int64_t)HUB_RUN, N - i0);
    int32_t my_c = -1, my_pt = 0;
    if (lane < cnt) { my_c = cand[i0 + lane]; my_pt = perm[i0 + lane]; }
    int cur = -2; float4 h[3]; float off[3] = {0.f, 0.f, 0.f};
    h[0] = h[1] = h[2] = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t bits = 0;
    for (int q = 0; q < cnt; ++q) {
        const int c = __shfl_sync(0xffffffffu, my_c, q);
        if (c < 0) continue;
        const int pt = __shfl_sync(0xffffffffu, my_pt, q);
        if (c != cur) {
            cur = c;
            for (int j = 0; j < 3; ++j) {
                h[j] = (ld <= 128 && lane * 4 < ld) ? *(const float4 *)(hp3 + ((size_t)c * 3 + j) * ld + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                off[j] = off3[c * 3 + j];
            }
        }
        const float *x = X + (int64_t)pt * ld;
        const int rel = (int)(i0 + q) - cstart[c];
        uint32_t s = 0;
        for (int j = 0; j < 3; ++j) {
            float margin = __fadd_rn(off[j], hub_lane_dot4(h[j], x, ld, lane, hp3 + ((size_t)c * 3 + j) * ld));
            uint32_t sj;
            if ((double)margin > 1e-8) sj = 0; else if ((double)margin < -1e-8) sj = 1; else sj = (uint32_t)(rel & 1);
            s |= sj << j;
        }
        bits |= s << (3 * (q & 7));   // 8 entries x 3 bits per 32-bit word half ... store per entry below
        if (lane == 0) side3[i0 + q] = (uint8_t)s;
    }
    (void)bits;
}

// one warp per candidate: count the left sides of the three candidates, pick the most balanced (:879-893), decide
// leaf vs split (:1080-1094 MIN_SPLIT_BALANCE 0.1), publish the chosen hyperplane and the partition flags
__global__ void k_hub_decide(int ld, int M, const int32_t *cstart, const int32_t *csize, const int32_t *cnode, const uint8_t *side3,
                             const float *hp3, const float *off3, int32_t *cleft, uint32_t *csplit, uint8_t *cchoice,
                             uint32_t *flag32, float *node_hyper, float *node_off, int64_t hyper_cap, int32_t *hyper_row_of_node,
                             int32_t *n_hyper, int *overflow) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + warp;
    if (c >= M) return;
    const int start = cstart[c], n = csize[c];
    int nl[3] = {0, 0, 0};
    for (int i = lane; i < n; i += 32) { uint8_t s = side3[start + i]; nl[0] += !(s & 1); nl[1] += !(s & 2); nl[2] += !(s & 4); }
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) nl[j] += __shfl_xor_sync(0xffffffffu, nl[j], o);
    float best = 0.0f; int choice = -1;
    for (int j = 0; j < 3; ++j) {
        if (isnan(off3[c * 3 + j])) continue;
        int a = nl[j], b = n - nl[j];
        if (a == 0 || b == 0) continue;
        float bal = __fdiv_rn((float)(a < b ? a : b), (float)n);
        if (bal > best) { best = bal; choice = j; }
    }
    const bool split = choice >= 0 && !(best < 0.1f);
    if (lane == 0) { csplit[c] = split; cchoice[c] = (uint8_t)(choice < 0 ? 0 : choice); cleft[c] = split ? nl[choice] : 0; }
    int row = -1;
    if (split) {
        if (lane == 0) { row = atomicAdd(n_hyper, 1); if (row >= hyper_cap) { *overflow = 1; row = -1; } hyper_row_of_node[cnode[c]] = row; if (row >= 0) node_off[row] = off3[c * 3 + choice]; }
        row = __shfl_sync(0xffffffffu, row, 0);
        if (row >= 0) for (int e = lane; e < ld; e += 32) node_hyper[(size_t)row * ld + e] = hp3[((size_t)c * 3 + choice) * ld + e];
    }
    for (int i = lane; i < n; i += 32) flag32[start + i] = split ? (uint32_t)!((side3[start + i] >> choice) & 1) : 0u;
}

// children of the split candidates: node ids next_base + 2*rank (+1), node table entries, next level's candidate flags
__global__ void k_hub_children(int M, const int32_t *cstart, const int32_t *csize, const int32_t *cnode, const int32_t *cleft,
                               const uint32_t *csplit, const uint32_t *csplit_scan, int32_t next_base, int leaf_size, int depth_ok,
                               int32_t *node_start, int32_t *node_size, int32_t *node_left, int32_t *node_right, uint32_t *child_isc) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= M) return;
    if (!csplit[c]) return;
    const int r = (int)csplit_scan[c];
    const int l = next_base + 2 * r, rr = l + 1;
    const int start = cstart[c], n = csize[c], nl = cleft[c];
    node_left[cnode[c]] = l; node_right[cnode[c]] = rr;
    node_start[l] = start; node_size[l] = nl; node_left[l] = -1; node_right[l] = -1;
    node_start[rr] = start + nl; node_size[rr] = n - nl; node_left[rr] = -1; node_right[rr] = -1;
    child_isc[2 * r] = depth_ok && nl > leaf_size;
    child_isc[2 * r + 1] = depth_ok && (n - nl) > leaf_size;
}
__global__ void k_hub_next(int n_children, int32_t next_base, const uint32_t *child_isc, const uint32_t *child_scan,
                           const int32_t *node_start, const int32_t *node_size, int32_t *cstart2, int32_t *csize2, int32_t *cnode2) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_children || !child_isc[j]) return;
    int w = (int)child_scan[j];
    cstart2[w] = node_start[next_base + j]; csize2[w] = node_size[next_base + j]; cnode2[w] = next_base + j;
}
// stable partition of the split candidates' segments + next-level candidate ordinal per position
__global__ void k_hub_scatter(int64_t N, const int32_t *perm, const int32_t *cand, const uint8_t *side3, const uint8_t *cchoice,
                              const uint32_t *pl, const int32_t *cstart, const int32_t *csize, const int32_t *cleft,
                              const uint32_t *csplit, const uint32_t *csplit_scan, const uint32_t *child_isc,
                              const uint32_t *child_scan, int32_t *perm2, int32_t *cand2) {
    int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= N) return;
    const int c = cand[pos];
    if (c < 0 || !csplit[c]) { perm2[pos] = perm[pos]; cand2[pos] = -1; return; }
    const int start = cstart[c], nl = cleft[c];
    const int lrank = (int)(pl[pos] - pl[start]);
    const bool left = !((side3[pos] >> cchoice[c]) & 1);
    const int64_t np = left ? (int64_t)start + lrank : (int64_t)start + nl + ((pos - start) - lrank);
    const int child = 2 * (int)csplit_scan[c] + (left ? 0 : 1);
    perm2[np] = perm[pos];
    cand2[np] = child_isc[child] ? (int32_t)child_scan[child] : -1;
}

extern "C" int nnd_hub_tree_workspace_bytes(int64_t N, int32_t d, size_t *bytes) {
    NND_CHECK(bytes && N > 0 && d > 0, NND_ERR_INVALID, "nnd_hub_tree_workspace_bytes: bad args");
    const int ld = (d + 3) & ~3;
    WsAlloc A(nullptr, 0);
    A.get<int32_t>(N); A.get<int32_t>(N); A.get<int32_t>(N); A.get<int32_t>(N); A.get<int32_t>(N + 1);   // perm x2, cand x2, deg
    A.get<uint8_t>(N + 16); A.get<uint32_t>(N + 1); A.get<uint32_t>(N + 2); A.get<uint32_t>(nnd_scan_tmp_elems(N + 2));
    const int64_t MC = N / 2 + 64;
    for (int i = 0; i < 6; ++i) A.get<int32_t>(MC);      // cstart/csize/cnode x2
    A.get<int32_t>(MC); A.get<uint32_t>(MC + 1); A.get<uint32_t>(MC + 2); A.get<uint8_t>(MC + 16);
    A.get<uint32_t>(2 * MC + 2); A.get<uint32_t>(2 * MC + 4);
    A.get<float>((size_t)(N / 8 + 64) * 3 * ld); A.get<float>((size_t)(N / 8 + 64) * 3);
    A.get<int32_t>(16);
    *bytes = A.off + 8192;
    return NND_OK;
}

// Build the hub tree.  Outputs (device): indices[N] = points in leaf order (the reference's tree.indices after
// convert_tree_format == _vertex_order), node table arrays of capacity 2N (start,size,left,right,hyper_row), chosen
// hyperplanes node_hyper[hyper_cap][ld] + node_off.  n_nodes_host / n_hyper_host receive the counts; the caller
// flattens the node table to the reference's pre-order arrays (rp_trees.py:2926-2961) on the host.
extern "C" int nnd_hub_tree_build(nnd_ctx *ctx, const float *X, int64_t N, int32_t d, int32_t ld, int32_t angular,
                                  const int32_t *ng_idx, int32_t k, int32_t leaf_size, int32_t max_depth, int32_t *indices,
                                  int32_t *node_start, int32_t *node_size, int32_t *node_left, int32_t *node_right,
                                  int32_t *node_hyper_row, float *node_hyper, float *node_off, int64_t hyper_cap,
                                  int32_t *n_nodes_host, int32_t *n_hyper_host, void *ws, size_t ws_bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    NND_CHECK(ctx && X && ng_idx && indices && node_start && node_size && node_left && node_right && node_hyper_row && node_hyper &&
              node_off && n_nodes_host && n_hyper_host && ws, NND_ERR_INVALID, "nnd_hub_tree_build: null argument");
    NND_CHECK(N > 0 && N < (1LL << 30) && ld % 4 == 0 && ld >= d && leaf_size >= 1, NND_ERR_INVALID, "nnd_hub_tree_build: bad shape");
    WsAlloc A(ws, ws_bytes);
    int32_t *perm[2] = {A.get<int32_t>(N), A.get<int32_t>(N)}, *cand[2] = {A.get<int32_t>(N), A.get<int32_t>(N)};
    int32_t *deg = A.get<int32_t>(N + 1);
    uint8_t *side3 = A.get<uint8_t>(N + 16); uint32_t *flag32 = A.get<uint32_t>(N + 1), *pl = A.get<uint32_t>(N + 2);
    uint32_t *scan_tmp = A.get<uint32_t>(nnd_scan_tmp_elems(N + 2));
    const int64_t MC = N / 2 + 64;
    int32_t *cstart[2], *csize[2], *cnode[2];
    for (int i = 0; i < 2; ++i) { cstart[i] = A.get<int32_t>(MC); csize[i] = A.get<int32_t>(MC); cnode[i] = A.get<int32_t>(MC); }
    int32_t *cleft = A.get<int32_t>(MC); uint32_t *csplit = A.get<uint32_t>(MC + 1), *csplit_scan = A.get<uint32_t>(MC + 2);
    uint8_t *cchoice = A.get<uint8_t>(MC + 16);
    uint32_t *child_isc = A.get<uint32_t>(2 * MC + 2), *child_scan = A.get<uint32_t>(2 * MC + 4);
    const int64_t HP3 = N / 8 + 64;
    float *hp3 = A.get<float>((size_t)HP3 * 3 * ld), *off3 = A.get<float>((size_t)HP3 * 3);
    int32_t *scal = A.get<int32_t>(16);
    NND_CHECK(A.ok, NND_ERR_WORKSPACE, "nnd_hub_tree_build: workspace too small");

    NND_CUDA(cudaMemsetAsync(deg, 0, sizeof(int32_t) * (N + 1), st));
    NND_CUDA(cudaMemsetAsync(scal, 0, sizeof(int32_t) * 16, st));
    k_hub_degrees<<<(unsigned)cdiv(N * k, 256), 256, 0, st>>>(ng_idx, N, k, deg); NND_LAUNCH(ctx);
    const bool root_cand = N > leaf_size && max_depth > 0;
    k_hub_init<<<(unsigned)cdiv(N, 256), 256, 0, st>>>(perm[0], cand[0], N, root_cand ? 1 : 0); NND_LAUNCH(ctx);
    // node 0 = root
    int32_t h0[4] = {0, (int32_t)N, -1, -1};
    NND_CUDA(cudaMemcpyAsync(node_start, &h0[0], 4, cudaMemcpyHostToDevice, st));
    NND_CUDA(cudaMemcpyAsync(node_size, &h0[1], 4, cudaMemcpyHostToDevice, st));
    NND_CUDA(cudaMemcpyAsync(node_left, &h0[2], 4, cudaMemcpyHostToDevice, st));
    NND_CUDA(cudaMemcpyAsync(node_right, &h0[3], 4, cudaMemcpyHostToDevice, st));
    NND_CUDA(cudaMemsetAsync(node_hyper_row, 0xff, sizeof(int32_t) * 2 * N, st));
    int M = 0, cur = 0, cc = 0;
    int32_t n_nodes = 1;
    if (root_cand) {
        int32_t hc[3] = {0, (int32_t)N, 0};
        NND_CUDA(cudaMemcpyAsync(cstart[0], &hc[0], 4, cudaMemcpyHostToDevice, st));
        NND_CUDA(cudaMemcpyAsync(csize[0], &hc[1], 4, cudaMemcpyHostToDevice, st));
        NND_CUDA(cudaMemcpyAsync(cnode[0], &hc[2], 4, cudaMemcpyHostToDevice, st));
        NND_CUDA(cudaStreamSynchronize(st));
        M = 1;
    }
    for (int level = 0; M > 0; ++level) {
        NND_CHECK(M <= HP3, NND_ERR_OVERFLOW, "hub tree: %d candidate nodes exceed the hyperplane scratch", M);
        const int hw = 4;
        k_hub_top3<<<(unsigned)cdiv(M, hw), hw * 32, (size_t)hw * ld * 4, st>>>(X, ld, angular, M, cstart[cc], csize[cc], perm[cur], deg, hp3, off3);
        NND_LAUNCH(ctx);
        k_hub_side<<<(unsigned)cdiv(cdiv(N, HUB_RUN) * 32, 256), 256, 0, st>>>(X, ld, N, perm[cur], cand[cur], cstart[cc], hp3, off3, side3);
        NND_LAUNCH(ctx);
        NND_CUDA(cudaMemsetAsync(flag32, 0, sizeof(uint32_t) * (N + 1), st));
        k_hub_decide<<<(unsigned)cdiv(M, hw), hw * 32, 0, st>>>(ld, M, cstart[cc], csize[cc], cnode[cc], side3, hp3, off3, cleft, csplit, cchoice,
                                                              flag32, node_hyper, node_off, hyper_cap, node_hyper_row, scal, scal + 1);
        NND_LAUNCH(ctx);
        NND_TRY(nnd_exclusive_scan_u32(ctx, flag32, pl, N, scan_tmp, st));
        NND_TRY(nnd_exclusive_scan_u32(ctx, csplit, csplit_scan, M, scan_tmp, st));
        uint32_t n_split = 0;
        NND_CUDA(cudaMemcpyAsync(&n_split, csplit_scan + M, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        NND_CUDA(cudaStreamSynchronize(st));
        const int n_children = 2 * (int)n_split;
        NND_CHECK((int64_t)n_nodes + n_children <= 2 * N, NND_ERR_OVERFLOW, "hub tree: node table overflow");
        const int depth_ok = (max_depth - (level + 1)) > 0;
        uint32_t m_next = 0;
        if (n_children > 0) {
            k_hub_children<<<(unsigned)cdiv(M, 256), 256, 0, st>>>(M, cstart[cc], csize[cc], cnode[cc], cleft, csplit, csplit_scan, n_nodes,
                                                                  leaf_size, depth_ok, node_start, node_size, node_left, node_right, child_isc);
            NND_LAUNCH(ctx);
            NND_TRY(nnd_exclusive_scan_u32(ctx, child_isc, child_scan, n_children, scan_tmp, st));
            k_hub_next<<<(unsigned)cdiv(n_children, 256), 256, 0, st>>>(n_children, n_nodes, child_isc, child_scan, node_start, node_size,
                                                                       cstart[cc ^ 1], csize[cc ^ 1], cnode[cc ^ 1]);
            NND_LAUNCH(ctx);
            NND_CUDA(cudaMemcpyAsync(&m_next, child_scan + n_children, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        }
        k_hub_scatter<<<(unsigned)cdiv(N, 256), 256, 0, st>>>(N, perm[cur], cand[cur], side3, cchoice, pl, cstart[cc], csize[cc], cleft, csplit,
                                                             csplit_scan, child_isc, child_scan, perm[cur ^ 1], cand[cur ^ 1]);
        NND_LAUNCH(ctx);
        NND_CUDA(cudaStreamSynchronize(st));
        n_nodes += n_children;
        cur ^= 1; cc ^= 1; M = (int)m_next;
    }
    NND_CUDA(cudaMemcpyAsync(indices, perm[cur], sizeof(int32_t) * N, cudaMemcpyDeviceToDevice, st));
    int32_t hs[2];
    NND_CUDA(cudaMemcpyAsync(hs, scal, sizeof(hs), cudaMemcpyDeviceToHost, st));
    NND_CUDA(cudaStreamSynchronize(st));
    NND_CHECK(hs[1] == 0, NND_ERR_OVERFLOW, "hub tree: more than %lld internal nodes (hyper_cap)", (long long)hyper_cap);
    *n_nodes_host = n_nodes; *n_hyper_host = hs[0];
    NND_CUDA(cudaGetLastError());
    return NND_OK;
}

// ------------------------------------------------------------------ diversification
// One warp per row of a row-major neighbour table (stride k, ascending distance, -1 = end / hole): entry j is dropped
// if some KEPT earlier entry c has  dist[c] > FLOAT32_EPS  and  d(x_j, x_c) < dist[j]  (pynndescent_.py:369-403; the
// same rule is what diversify_csr :549-588 applies to rows already in ascending order).  Rows are rewritten
// compacted with (-1, +inf) padding.
__global__ void __launch_bounds__(128) k_diversify(const float *X, int64_t N, int d, int ld, int metric, const void *aux, int32_t *ind,
                                                   float *dist, int k) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (row >= N) return;
    int32_t *ri = (int32_t *)smem_raw + (size_t)warp * k * 4;   // [k] in idx, [k] kept idx
    float *rd = (float *)(ri + 2 * k);                            // [k] in dist, [k] kept dist
    int32_t *ki = ri + k; float *kd = rd + k;
    for (int j = lane; j < k; j += 32) { ri[j] = ind[row * k + j]; rd[j] = dist[row * k + j]; }
    __syncwarp();
    int nk = 0;
    if (lane == 0) { ki[0] = ri[0]; kd[0] = rd[0]; }
    nk = 1;
    __syncwarp();
    for (int j = 1; j < k; ++j) {
        const int32_t pj = ri[j];
        if (pj < 0) break;
        const float dj = rd[j];
        bool hit = false;
        for (int c = lane; c < nk; c += 32) {
            float dd = pair_distance_seq(X + (int64_t)pj * ld, X + (int64_t)ki[c] * ld, d, metric, aux, pj, ki[c]);
            hit |= (kd[c] > 1.1920929e-07f && dd < dj);
        }
        if (!__any_sync(0xffffffffu, hit)) { if (lane == 0) { ki[nk] = pj; kd[nk] = dj; } ++nk; }
        __syncwarp();
    }
    for (int j = lane; j < k; j += 32) {
        ind[row * k + j] = j < nk ? ki[j] : -1;
        dist[row * k + j] = j < nk ? kd[j] : CUDART_INF_F;
    }
}
extern "C" int nnd_diversify(nnd_ctx *ctx, const nnd_problem *P, int32_t *ind, float *dist, void *stream) {
    NND_CHECK(ctx && P && ind && dist, NND_ERR_INVALID, "nnd_diversify: bad args");
    const int warps = 4;
    size_t smem = (size_t)warps * P->k * 16;
    NND_CHECK(smem <= 96 * 1024, NND_ERR_UNSUPPORTED, "nnd_diversify: k=%d too large", P->k);
    k_diversify<<<(unsigned)cdiv(P->N, warps), warps * 32, smem, (cudaStream_t)stream>>>(P->X, P->N, P->d, P->ld, P->metric, P->aux, ind, dist, P->k);
    NND_LAUNCH(ctx);
    NND_CUDA(cudaGetLastError());
    return NND_OK;
}
