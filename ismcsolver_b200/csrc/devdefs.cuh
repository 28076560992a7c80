// devdefs.cuh — compiler shims shared by all code of the engine.
//
// The games and the search body are written once as inline functions.  Under
// nvcc they are __host__ __device__: the kernels (engine.cu) run them on the
// GPU, and the host game helpers of the C ABI (games_host.cu: deal / legal /
// apply / result, no search) run the very same rules on the host.  The same
// headers also compile with plain g++; tests/emu uses that to check the device
// logic against the oracle on a machine without a GPU (test infrastructure, not
// part of the product library — the search itself has no host path).
#pragma once

#include <stdint.h>
#include <chrono>

#if defined(__CUDACC__)
#define ISMCTS_DEV __host__ __device__ __forceinline__
#define ISMCTS_HDI __host__ __device__ __forceinline__
// a real call: keeps the registers of rarely executed code out of the hot loops (lane_search.cuh)
#define ISMCTS_DEV_CALL __host__ __device__ __noinline__
#else
#define ISMCTS_DEV inline
#define ISMCTS_HDI inline
#define ISMCTS_DEV_CALL inline
#endif

#include "../../include/ismcts_detmath.h"

namespace ismcts {

#if defined(__CUDA_ARCH__)
ISMCTS_DEV uint32_t popc32(uint32_t x) { return (uint32_t)__popc(x); }
ISMCTS_DEV uint32_t popc64(uint64_t x) { return (uint32_t)__popcll(x); }
ISMCTS_DEV uint32_t ctz32(uint32_t x) { return (uint32_t)(__ffs((int)x) - 1); }
ISMCTS_DEV uint32_t clz32(uint32_t x) { return (uint32_t)__clz((int)x); }
ISMCTS_DEV uint32_t ctz64(uint64_t x) { return (uint32_t)(__ffsll((long long)x) - 1); }
ISMCTS_DEV uint32_t mulhi32(uint32_t a, uint32_t b) { return __umulhi(a, b); }
ISMCTS_DEV uint64_t global_timer_ns()
{
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Atomics on the node pool and the result arrays.  All of them name the GLOBAL state space: through a generic pointer
// the compiler adds a run-time test for shared memory and a compare-and-swap loop for that case to every 64-bit
// atomic, and turns an atomicAdd() whose result is dropped into an ATOM rather than a fire-and-forget RED.
ISMCTS_DEV void atomic_add_u64(uint64_t* p, uint64_t v) { asm volatile("red.global.add.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
// Warp vote over all 32 lanes (every lane of a warp runs the search loop, see lane_search.cuh).
ISMCTS_DEV bool warp_any(bool p) { return __any_sync(0xFFFFFFFFu, p) != 0; }
// Shared-tree (tree-parallel) primitives: slot reservation, child publication, counters.
ISMCTS_DEV uint32_t atomic_cas_u32(uint32_t* p, uint32_t expect, uint32_t value)
{
    uint32_t old;
    asm volatile("atom.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(p), "r"(expect), "r"(value) : "memory");
    return old;
}
ISMCTS_DEV uint32_t atomic_add_u32(uint32_t* p, uint32_t v)
{
    uint32_t old;
    asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
ISMCTS_DEV void red_add_u32(uint32_t* p, uint32_t v) { asm volatile("red.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
// Publishes: everything this thread wrote before is visible to a thread that sees the bit (release at GPU scope — a
// MEMBAR, unlike __threadfence() no invalidation of the SM's whole L1, which every lane's local memory lives in).
ISMCTS_DEV void atomic_or_u64_release(uint64_t* p, uint64_t v) { asm volatile("red.release.gpu.global.or.b64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
ISMCTS_DEV void atomic_add_f64(double* p, double v) { asm volatile("red.global.add.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory"); }
// (the rare path of a lane that waits for another lane's publication)
ISMCTS_DEV uint64_t load_u64_acquire(const uint64_t* p)
{
    uint64_t v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// Starts bringing the line at p into L2 (no register, no wait).
ISMCTS_DEV void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
// A table in global memory nobody writes during the kernel (ln): the read-only path through L1.
ISMCTS_DEV double load_f64_readonly(const double* p) { return __ldg(p); }
template <class T> ISMCTS_DEV T load_factors_readonly(const T* p)
{
    const double2 v = __ldg(reinterpret_cast<const double2*>(p));
    return T{v.x, v.y};
}
ISMCTS_DEV uint32_t load_u32_cg(const uint32_t* p) { return __ldcg(p); }
ISMCTS_DEV uint64_t load_u64_cg(const uint64_t* p) { return __ldcg((const unsigned long long*)p); }
ISMCTS_DEV double load_f64_cg(const double* p) { return __ldcg(p); }
// Fire-and-forget 64-bit reduction (SASS RED.E.ADD.64): no return value, no scoreboard wait.
ISMCTS_DEV void red_add_u64(uint64_t* p, uint64_t v)
{
    asm volatile("red.global.add.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
#else
ISMCTS_DEV uint32_t popc32(uint32_t x) { return (uint32_t)__builtin_popcount(x); }
ISMCTS_DEV uint32_t popc64(uint64_t x) { return (uint32_t)__builtin_popcountll(x); }
ISMCTS_DEV uint32_t ctz32(uint32_t x) { return (uint32_t)__builtin_ctz(x); }
ISMCTS_DEV uint32_t clz32(uint32_t x) { return x ? (uint32_t)__builtin_clz(x) : 32u; }
ISMCTS_DEV uint32_t ctz64(uint64_t x) { return (uint32_t)__builtin_ctzll(x); }
ISMCTS_DEV uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
ISMCTS_DEV uint64_t global_timer_ns()
{
    return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(
               std::chrono::steady_clock::now().time_since_epoch()).count();
}
// host build: a "warp" of one lane, unless a test harness that runs several lanes (tests/emu:
// fibers) installs its vote
inline bool (*g_host_warp_any)(bool) = nullptr;
ISMCTS_DEV bool warp_any(bool p) { return g_host_warp_any ? g_host_warp_any(p) : p; }
ISMCTS_DEV uint32_t atomic_cas_u32(uint32_t* p, uint32_t expect, uint32_t value)
{
    const uint32_t old = *p;
    if (old == expect) *p = value;
    return old;
}
ISMCTS_DEV uint32_t atomic_add_u32(uint32_t* p, uint32_t v) { const uint32_t old = *p; *p += v; return old; }
ISMCTS_DEV void red_add_u32(uint32_t* p, uint32_t v) { *p += v; }
ISMCTS_DEV void atomic_or_u64_release(uint64_t* p, uint64_t v) { *p |= v; }
ISMCTS_DEV void atomic_add_f64(double* p, double v) { *p = ISMCTS_DADD(*p, v); }
ISMCTS_DEV uint64_t load_u64_acquire(const uint64_t* p) { return *p; }
ISMCTS_DEV void prefetch_l2(const void*) {}
ISMCTS_DEV double load_f64_readonly(const double* p) { return *p; }
template <class T> ISMCTS_DEV T load_factors_readonly(const T* p) { return *p; }
ISMCTS_DEV uint32_t load_u32_cg(const uint32_t* p) { return *p; }
ISMCTS_DEV uint64_t load_u64_cg(const uint64_t* p) { return *p; }
ISMCTS_DEV double load_f64_cg(const double* p) { return *p; }
ISMCTS_DEV void atomic_add_u64(uint64_t* p, uint64_t v) { *p += v; }
ISMCTS_DEV void red_add_u64(uint64_t* p, uint64_t v) { *p += v; }
#endif

// ---- the scratch of games that find it themselves (Knockout Whist, kw_game.cuh)
//
// Such a game holds no pointer to its scratch: it is the column of the thread in the block's dynamic shared
// memory, found from the thread index alone.  Host builds (the C ABI's game helpers, tests/emu) install the
// scratch of the "block" and the index of the thread that is running before they call game code.
constexpr uint32_t kWarpLanes = 32;
constexpr uint32_t kBlockLanes = 128;      // threads per block of every kernel that runs games (engine.cu: kBlock)
inline thread_local uint64_t* g_host_block_scratch = nullptr;
inline thread_local uint32_t g_host_block_lane = 0;
#if defined(__CUDA_ARCH__)
ISMCTS_DEV uint64_t* block_scratch()
{
    extern __shared__ uint64_t ismcts_dynamic_smem[];
    return ismcts_dynamic_smem;
}
ISMCTS_DEV uint32_t block_lane() { return threadIdx.x; }
ISMCTS_DEV void warp_sync() { __syncwarp(); }
#else
ISMCTS_DEV uint64_t* block_scratch() { return g_host_block_scratch; }
ISMCTS_DEV uint32_t block_lane() { return g_host_block_lane; }
// (tests/emu: a vote is the only point where its fibers switch, so a warp_sync() must be one)
ISMCTS_DEV void warp_sync() { (void)warp_any(false); }
#endif

// Index of the k-th (0-based) set bit of w; k < popc(w).
ISMCTS_DEV uint32_t kth_bit32(uint32_t w, uint32_t k)
{
    uint32_t pos = 0, c;
    c = popc32(w & 0xFFFFu); if (k >= c) { k -= c; pos = 16; w >>= 16; }
    c = popc32(w & 0xFFu);   if (k >= c) { k -= c; pos += 8; w >>= 8; }
    c = popc32(w & 0xFu);    if (k >= c) { k -= c; pos += 4; w >>= 4; }
    c = popc32(w & 0x3u);    if (k >= c) { k -= c; pos += 2; w >>= 2; }
    c = w & 1u;              if (k >= c) { pos += 1; }
    return pos;
}

ISMCTS_DEV uint32_t kth_bit64(uint64_t m, uint32_t k)
{
    // choose the half first so that the 32-bit search exists once in the code
    const uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);
    const uint32_t c = popc32(lo);
    const bool upper = k >= c;
    return (upper ? 32u : 0u) + kth_bit32(upper ? hi : lo, upper ? k - c : k);
}

} // namespace ismcts
