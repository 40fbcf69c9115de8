// cuda_emu.h -- TEST INFRASTRUCTURE: a minimal host emulation of the CUDA execution model, just large enough to run the
// SOURCE of selected kernels of libnnd_b200 on the CPU (tests/test_kernel_emulation.py extracts the kernels textually from
// pynndescent_b200/csrc/*.cu and compiles them against this header with g++).
//
// Model: one OS thread per CUDA thread, blocks executed one after the other.  __syncthreads / __syncwarp / shuffles /
// ballots are real barriers between the participating threads (all kernels emulated here call them warp- or block-
// uniformly with full masks), shared memory is a per-block buffer, `__shared__` statics are plain statics (blocks are
// sequential), mbarriers follow the PTX semantics (phase completes when pending arrivals and pending transaction bytes
// are both zero; try_wait.parity succeeds when the phase of that parity has completed) and cp.async.bulk is served by an
// asynchronous "TMA" thread per block that copies and then signals complete_tx -- so the producer/consumer protocol of
// a warp-specialised kernel runs with genuine concurrency.  Arithmetic intrinsics map to IEEE operations (compile with
// -ffp-contract=off), so results are bit-comparable with the oracle.
#pragma once
#include <atomic>
#include <chrono>
#include <barrier>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <limits>
#include <memory>
#include <mutex>
#include <thread>
#include <type_traits>
#include <unordered_map>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __shared__ static
#define CUDART_INF_F (std::numeric_limits<float>::infinity())
#define CUDART_INF (std::numeric_limits<double>::infinity())
#define CUDART_NAN_F (std::numeric_limits<float>::quiet_NaN())
using std::isnan;

typedef void *cudaStream_t;
typedef void *cudaEvent_t;
typedef int cudaError_t;
static const int cudaSuccess = 0;
static inline const char *cudaGetErrorString(int) { return "emulated"; }

struct float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct uint4 { unsigned x, y, z, w; };
struct dim3 { unsigned x = 1, y = 1, z = 1; };
static thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

namespace emu {
// EMU_CHAOS=<permille>: every synchronisation primitive first sleeps for a random 0-300 us with that probability, so that
// repeated runs visit very different interleavings of producer lanes, consumer warps and the copy engine
static inline void chaos() {
    static const int permille = [] { const char *e = getenv("EMU_CHAOS"); return e ? atoi(e) : 0; }();
    if (permille <= 0) return;
    static thread_local unsigned s = (unsigned)std::hash<std::thread::id>()(std::this_thread::get_id()) | 1u;
    s = s * 1664525u + 1013904223u;
    if ((int)((s >> 10) % 1000u) < permille) std::this_thread::sleep_for(std::chrono::microseconds((s >> 20) % 300u));
}
struct MBar {
    std::mutex m; std::condition_variable cv;
    int count = 0, pending = 0; long tx = 0; unsigned phase = 0;
    void maybe_complete() { if (pending == 0 && tx == 0) { ++phase; pending = count; cv.notify_all(); } }
    void arrive(long expect) {
        std::lock_guard<std::mutex> g(m);
        if (pending <= 0) { fprintf(stderr, "emu: mbarrier received more arrivals than expected\n"); abort(); }
        tx += expect; --pending; maybe_complete();
    }
    void complete_tx(long n) { std::lock_guard<std::mutex> g(m); tx -= n; maybe_complete(); }
    void wait(unsigned parity);
};
struct Warp {
    std::barrier<> bar{32}; std::barrier<> hbar0{16}, hbar1{16}; uint64_t slot[32];
    // full-warp masks and the two half-warp masks are supported (half-warp collectives)
    std::barrier<> &group(unsigned mask) {
        if (mask == 0xffffffffu) return bar;
        if (mask == 0x0000ffffu) return hbar0;
        if (mask == 0xffff0000u) return hbar1;
        fprintf(stderr, "emu: unsupported lane mask %08x\n", mask); abort();
    }
};
struct Copy { void *dst; const void *src; size_t bytes; MBar *bar; };
struct Block {
    int nthreads;
    std::unique_ptr<unsigned char[]> smem_store; unsigned char *smem = nullptr;
    std::barrier<> bar;
    std::vector<std::unique_ptr<Warp>> warps;
    std::mutex bm; std::unordered_map<const void *, std::unique_ptr<MBar>> bars;
    std::mutex qm; std::condition_variable qcv; std::deque<Copy> q; bool stop = false; std::thread tma; unsigned lcg = 12345;
    Block(int n, size_t smem_bytes) : nthreads(n), bar(n) {
        smem_store.reset(new unsigned char[smem_bytes + 256]);
        smem = (unsigned char *)(((uintptr_t)smem_store.get() + 127) & ~uintptr_t(127));
        for (int w = 0; w < (n + 31) / 32; ++w) warps.emplace_back(new Warp());
        tma = std::thread([this] {
            for (;;) {
                Copy c;
                {
                    std::unique_lock<std::mutex> g(qm);
                    qcv.wait(g, [&] { return stop || !q.empty(); });
                    if (q.empty()) return;
                    lcg = lcg * 1664525u + 1013904223u;                 // serve the pending copies in a scrambled order
                    size_t pick = (lcg >> 8) % q.size();
                    c = q[pick]; q.erase(q.begin() + pick);
                }
                chaos();
                memcpy(c.dst, c.src, c.bytes);
                c.bar->complete_tx((long)c.bytes);
            }
        });
    }
    ~Block() { { std::lock_guard<std::mutex> g(qm); stop = true; } qcv.notify_all(); tma.join(); }
    MBar *mbar(const void *p) {
        std::lock_guard<std::mutex> g(bm);
        auto it = bars.find(p);
        if (it == bars.end()) { fprintf(stderr, "emu: mbarrier used before init\n"); abort(); }
        return it->second.get();
    }
};
static Block *g_block = nullptr;
static thread_local int t_tid = 0;
// a wait that lasts 4 minutes is a deadlock: print who waits on what, then abort (the test fails instead of hanging)
inline void MBar::wait(unsigned parity) {
    std::unique_lock<std::mutex> g(m);
    for (int spins = 0; !cv.wait_for(g, std::chrono::seconds(1), [&] { return (phase & 1u) != parity; }); ++spins) {
        if (spins < 240) continue;     // generous: the suite may share a loaded machine with other emulations
        long off = -1;
        for (auto &kv : g_block->bars) if (kv.second.get() == this) off = (long)((const unsigned char *)kv.first - g_block->smem);
        fprintf(stderr, "emu: DEADLOCK block %u thread %d waits parity %u on mbarrier@smem+%ld (count %d pending %d tx %ld phase %u)\n",
                blockIdx.x, t_tid, parity, off, count, pending, tx, phase);
        for (auto &kv : g_block->bars)
            fprintf(stderr, "     bar@%ld count %d pending %d tx %ld phase %u\n", (long)((const unsigned char *)kv.first - g_block->smem),
                    kv.second->count, kv.second->pending, kv.second->tx, kv.second->phase);
        abort();
    }
}
static inline Warp &warp() { return *g_block->warps[t_tid >> 5]; }
static inline int lane() { return t_tid & 31; }
static inline unsigned char *dyn_smem() { return g_block->smem; }

// run `grid` blocks of `block` threads one after the other; every CUDA thread is an OS thread
template <class F>
static void launch(int grid, int block, size_t smem_bytes, F kernel) {
    for (int b = 0; b < grid; ++b) {
        Block B(block, smem_bytes);
        g_block = &B;
        std::vector<std::thread> th;
        for (int t = 0; t < block; ++t)
            th.emplace_back([&, t] {
                t_tid = t;
                threadIdx.x = (unsigned)t; blockIdx.x = (unsigned)b; blockDim.x = (unsigned)block; gridDim.x = (unsigned)grid;
                kernel();
            });
        for (auto &x : th) x.join();
        g_block = nullptr;
    }
}
// kernels without any thread cooperation: all threads of the grid run sequentially on the calling thread
template <class F>
static void launch_seq(long grid, int block, F kernel) {
    for (long b = 0; b < grid; ++b)
        for (int t = 0; t < block; ++t) {
            t_tid = t;
            threadIdx.x = (unsigned)t; blockIdx.x = (unsigned)b; blockDim.x = (unsigned)block; gridDim.x = (unsigned)grid;
            kernel();
        }
}
}  // namespace emu

// ---- synchronisation
static inline void __syncthreads() { emu::g_block->bar.arrive_and_wait(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::chaos(); emu::warp().group(mask).arrive_and_wait(); }
template <class T, class SEL>
static inline T emu_exchange(unsigned mask, T v, SEL pick) {
    static_assert(sizeof(T) <= 8, "shuffle of a type wider than 8 bytes");
    emu::Warp &w = emu::warp();
    std::barrier<> &b = w.group(mask);
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    w.slot[emu::lane()] = raw;
    b.arrive_and_wait();
    int src = pick(emu::lane());
    T r = v;
    if (src >= 0 && src < 32 && ((mask >> src) & 1u)) { uint64_t x = w.slot[src]; memcpy(&r, &x, sizeof(T)); }
    b.arrive_and_wait();
    return r;
}
template <class T> static inline T __shfl_sync(unsigned m, T v, int src) { return emu_exchange(m, v, [=](int) { return src & 31; }); }
template <class T> static inline T __shfl_up_sync(unsigned m, T v, int d) { return emu_exchange(m, v, [=](int l) { return l - d; }); }
template <class T> static inline T __shfl_xor_sync(unsigned m, T v, int x) { return emu_exchange(m, v, [=](int l) { return l ^ x; }); }
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    emu::Warp &w = emu::warp();
    std::barrier<> &b = w.group(mask);
    w.slot[emu::lane()] = pred ? 1u : 0u;
    b.arrive_and_wait();
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) if ((mask >> i) & 1u) r |= (unsigned)(w.slot[i] & 1u) << i;
    b.arrive_and_wait();
    return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
template <class T> static inline T __ldg(const T *p) { return *p; }
// lanes of `mask` holding the same 32-bit value
static inline unsigned __match_any_sync(unsigned mask, int v) {
    emu::Warp &w = emu::warp();
    std::barrier<> &b = w.group(mask);
    w.slot[emu::lane()] = (unsigned)v;
    b.arrive_and_wait();
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) if (((mask >> i) & 1u) && w.slot[i] == (unsigned)v) r |= 1u << i;
    b.arrive_and_wait();
    return r;
}

// ---- mbarrier + bulk copy (nnd_common.cuh's PTX wrappers, mbar_arrive)
static inline void mbar_init(uint64_t *bar, uint32_t count) {
    std::lock_guard<std::mutex> g(emu::g_block->bm);
    auto &p = emu::g_block->bars[bar];
    p.reset(new emu::MBar()); p->count = (int)count; p->pending = (int)count;
}
static inline void mbar_fence_init() {}
static inline void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { emu::chaos(); emu::g_block->mbar(bar)->arrive((long)bytes); }
static inline void mbar_arrive(uint64_t *bar) { emu::chaos(); emu::g_block->mbar(bar)->arrive(0); }
static inline void mbar_wait(uint64_t *bar, uint32_t parity) { emu::chaos(); emu::g_block->mbar(bar)->wait(parity); emu::chaos(); }
static inline void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    if (((uintptr_t)dst & 15) || ((uintptr_t)src & 15) || (bytes & 15)) { fprintf(stderr, "emu: misaligned bulk copy\n"); abort(); }
    emu::chaos();
    emu::Block *B = emu::g_block;
    emu::MBar *mb = B->mbar(bar);
    { std::lock_guard<std::mutex> g(B->qm); B->q.push_back(emu::Copy{dst, src, bytes, mb}); }
    B->qcv.notify_one();
}

// ---- atomics (global memory)
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicMin(unsigned *p, unsigned v) {
    unsigned old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}

static inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicAnd(unsigned *p, unsigned v) { return __atomic_fetch_and(p, v, __ATOMIC_RELAXED); }
static inline int atomicCAS(int *p, int cmp, int v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED); return cmp; }
static inline int atomicMin(int *p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
static inline int atomicMax(int *p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v > old && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}

// ---- arithmetic intrinsics (IEEE round-to-nearest; build with -ffp-contract=off)
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return sqrt(a); }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline long long __double_as_longlong(double x) { long long r; memcpy(&r, &x, 8); return r; }
static inline double __longlong_as_double(long long x) { double r; memcpy(&r, &x, 8); return r; }
static inline unsigned __float_as_uint(float x) { unsigned r; memcpy(&r, &x, 4); return r; }
static inline float __uint_as_float(unsigned x) { float r; memcpy(&r, &x, 4); return r; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
template <class A, class B> static inline typename std::common_type<A, B>::type min(A a, B b) { return a < b ? a : b; }
template <class A, class B> static inline typename std::common_type<A, B>::type max(A a, B b) { return a > b ? a : b; }
