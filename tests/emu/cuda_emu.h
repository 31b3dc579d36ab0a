// cuda_emu.h — just enough of the CUDA execution model to run polypolish_b200/csrc/polish_dev.cuh on the CPU (g++), so that
// the kernels' logic can be checked against the oracle without a GPU (tests/test_emu_polish.py).  TEST INFRASTRUCTURE: one OS
// thread per CUDA thread of a block, blocks one after another, __syncthreads = a block barrier, warp shuffles / votes through a
// per-warp exchange buffer + warp barrier.  Slow and simple; not part of the product.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define PP_EMULATE 1
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __grid_constant__
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __constant__ static

struct dim3 { unsigned x = 1, y = 1, z = 1; };
struct uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };
struct alignas(16) double2 { double x, y; };
struct alignas(32) double4 { double x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return {x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return {x, y, z, w}; }
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return {x, y}; }
static inline double4 make_double4(double x, double y, double z, double w) { return {x, y, z, w}; }

namespace emu {
struct Warp {
    std::barrier<> bar;
    unsigned long long xchg[32];
    explicit Warp(int n) : bar(n) {}
};
struct Block {
    std::barrier<> bar;
    std::vector<std::unique_ptr<Warp>> warps;
    std::vector<unsigned char> shared;
    size_t shared_used = 0;
    explicit Block(int n) : bar(n) {}
};
struct ThreadState { dim3 tid, bid, bdim, gdim; Block* block = nullptr; Warp* warp = nullptr; };
inline thread_local ThreadState T;

// Runs `body` as a grid of `grid` blocks of `threads` threads (1-D), with `shared_bytes` of block-shared memory.
inline void launch(unsigned grid, unsigned threads, size_t shared_bytes, const std::function<void()>& body) {
    for (unsigned b = 0; b < grid; ++b) {
        Block blk((int)threads);
        blk.shared.assign(shared_bytes + 64, 0xCD);          // uninitialised shared memory is garbage, like on the device
        for (unsigned w = 0; w < (threads + 31) / 32; ++w) blk.warps.emplace_back(new Warp((int)std::min(32u, threads - 32 * w)));
        std::vector<std::thread> th;
        for (unsigned t = 0; t < threads; ++t)
            th.emplace_back([&, t] {
                T.tid = {t, 0, 0}; T.bid = {b, 0, 0}; T.bdim = {threads, 1, 1}; T.gdim = {grid, 1, 1};
                T.block = &blk; T.warp = blk.warps[t / 32].get();
                body();
            });
        for (auto& x : th) x.join();
    }
}
inline void* shared_base() { return (void*)(((uintptr_t)T.block->shared.data() + 63) & ~(uintptr_t)63); }
}  // namespace emu

#define threadIdx (emu::T.tid)
#define blockIdx (emu::T.bid)
#define blockDim (emu::T.bdim)
#define gridDim (emu::T.gdim)

static inline void __syncthreads() { emu::T.block->bar.arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::T.warp->bar.arrive_and_wait(); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }

template <class V> static inline V emu_exchange(V v, int src_lane) {   // every thread of the warp calls it
    static_assert(sizeof(V) <= 8, "shuffle of <= 8 bytes");
    emu::Warp* w = emu::T.warp;
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof v);
    w->xchg[threadIdx.x & 31] = raw;
    w->bar.arrive_and_wait();
    const unsigned long long got = w->xchg[src_lane & 31];
    w->bar.arrive_and_wait();
    V out;
    memcpy(&out, &got, sizeof out);
    return out;
}
template <class V> static inline V __shfl_sync(unsigned, V v, int src) { return emu_exchange(v, src); }
template <class V> static inline V __shfl_up_sync(unsigned, V v, unsigned d) {
    const int lane = threadIdx.x & 31;
    V got = emu_exchange(v, lane >= (int)d ? lane - (int)d : lane);
    return lane >= (int)d ? got : v;
}
template <class V> static inline V __shfl_down_sync(unsigned, V v, unsigned d) {
    const int lane = threadIdx.x & 31;
    const int n = std::min(32u, blockDim.x - (threadIdx.x & ~31u));
    V got = emu_exchange(v, lane + (int)d < n ? lane + (int)d : lane);
    return lane + (int)d < n ? got : v;
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    emu::Warp* w = emu::T.warp;
    w->xchg[threadIdx.x & 31] = pred ? 1 : 0;
    w->bar.arrive_and_wait();
    unsigned m = 0;
    const int n = std::min(32u, blockDim.x - (threadIdx.x & ~31u));
    for (int i = 0; i < n; ++i) m |= (unsigned)w->xchg[i] << i;
    w->bar.arrive_and_wait();
    return m;
}
static inline unsigned __reduce_max_sync(unsigned, unsigned v) {
    unsigned m = v;
    for (int o = 16; o > 0; o >>= 1) m = std::max(m, __shfl_down_sync(0xffffffffu, m, o));
    return __shfl_sync(0xffffffffu, m, 0);
}

// ---- atomics (relaxed is what the device gives; the barriers order everything else)
template <class V> static inline V atomicAdd(V* p, V v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline double atomicAdd(double* p, double v) {
    unsigned long long* q = (unsigned long long*)p;
    unsigned long long old = __atomic_load_n(q, __ATOMIC_RELAXED), want;
    double od;
    do { memcpy(&od, &old, 8); const double nd = od + v; memcpy(&want, &nd, 8); } while (!__atomic_compare_exchange_n(q, &old, want, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    return od;
}
template <class V> static inline V atomicOr(V* p, V v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
template <class V> static inline V atomicMin(V* p, V v) {
    V old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
template <class V> static inline V atomicMax(V* p, V v) {
    V old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
template <class V> static inline V atomicCAS(V* p, V cmp, V val) {
    __atomic_compare_exchange_n(p, &cmp, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}
template <class V> static inline V __ldg(const V* p) { return *p; }
template <class V> static inline V __ldcg(const V* p) { return __atomic_load_n(p, __ATOMIC_SEQ_CST); }

// ---- intrinsics
static inline unsigned __brev(unsigned x) {
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    return __builtin_bswap32(x);
}
static inline unsigned long long __brevll(unsigned long long x) { return ((unsigned long long)__brev((unsigned)x) << 32) | __brev((unsigned)(x >> 32)); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) { return (unsigned)((((unsigned long long)hi << 32) | lo) >> (shift & 31)); }
static inline long long __double_as_longlong(double x) { long long r; memcpy(&r, &x, 8); return r; }
static inline double __longlong_as_double(long long x) { double r; memcpy(&r, &x, 8); return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
using std::max;
using std::min;
static inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
