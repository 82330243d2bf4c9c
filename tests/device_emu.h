// device_emu.h -- just enough of the CUDA execution model to run this repository's kernel bodies on the
// host, for tests (tests/lc_device_probe.cpp, tests/device_step_probe.cpp).  TEST INFRASTRUCTURE ONLY.
//
// A "block" is one warp: 32 lanes as cooperative fibers (ucontext) on ONE OS thread, switched round-robin at
// every warp-level primitive, so the lanes advance in lock step exactly where the kernels require it
// (__shfl*_sync, __ballot_sync, __syncwarp, __syncthreads) and run to completion in between.  blockDim.x is
// 32, so __syncthreads() is the same barrier and __shared__ is a static.  Atomics are plain operations (one
// OS thread).  Kernel bodies in this repository take (View, blockIndex, gridSize) and loop grid-stride, so a
// launch is: for every block, run the 32 fibers to the end.
#pragma once
#include <ucontext.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#include <vector_types.h>
#include <vector_functions.h>
#undef __device__
#undef __global__
#undef __forceinline__
#undef __launch_bounds__
#undef __align__
#undef __shared__
#define __device__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static

static struct { unsigned x = 32, y = 1, z = 1; } blockDim;
static struct { unsigned x = 0, y = 0, z = 0; } threadIdx;   // set by the scheduler on every fiber switch

namespace emu {

constexpr int LANES = 32;
struct Fiber {
    ucontext_t ctx;
    std::vector<char> stack;
    bool done = true;
};
static Fiber g_fiber[LANES];
static ucontext_t g_main;
static int g_cur = -1;
static unsigned g_barrierCount = 0, g_barrierGen = 0;
static uint64_t g_xchg[LANES];
static std::function<void()> g_body;

static void switchTo(int next) {
    const int prev = g_cur;
    g_cur = next;
    threadIdx.x = (unsigned) (next < 0 ? 0 : next);
    swapcontext(prev < 0 ? &g_main : &g_fiber[prev].ctx, next < 0 ? &g_main : &g_fiber[next].ctx);
}
static void yieldLane() {   // to the next lane that is still running (or back to main when none is)
    for (int k = 1; k <= LANES; ++k) {
        const int n = (g_cur + k) % LANES;
        if (!g_fiber[n].done) { if (n != g_cur) switchTo(n); return; }
    }
}
static void trampoline() {
    g_body();
    g_fiber[g_cur].done = true;
    for (int k = 1; k < LANES; ++k) {
        const int n = (g_cur + k) % LANES;
        if (!g_fiber[n].done) { switchTo(n); }
    }
    switchTo(-1);
}
// all 32 lanes of the block meet here
static void barrier() {
    const unsigned gen = g_barrierGen;
    if (++g_barrierCount == LANES) { g_barrierCount = 0; ++g_barrierGen; return; }
    while (g_barrierGen == gen) {
        const int before = g_cur;
        yieldLane();
        if (g_cur == before && g_barrierGen == gen) {   // nobody else can run: a lane left the kernel before a full-mask barrier
            fprintf(stderr, "emu: barrier deadlock (lane %d waits, others finished)\n", g_cur);
            abort();
        }
    }
}
// run `body` as one block of 32 lanes
static void runBlock(const std::function<void()> &body) {
    g_body = body;
    g_barrierCount = 0;
    for (int i = 0; i < LANES; ++i) {
        Fiber &f = g_fiber[i];
        if (f.stack.empty()) f.stack.resize(512 * 1024);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack.data();
        f.ctx.uc_stack.ss_size = f.stack.size();
        f.ctx.uc_link = nullptr;
        makecontext(&f.ctx, trampoline, 0);
        f.done = false;
    }
    switchTo(0);
    g_cur = -1;
}
// kernel bodies: f(blockIndex, gridSize)
template <class F>
static void launch(int nBlocks, F f) {
    for (int b = 0; b < nBlocks; ++b) runBlock([&, b]() { f(b, nBlocks); });
}

template <class T>
static T exchange(T v, int src) {
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    g_xchg[threadIdx.x] = bits;
    barrier();
    const uint64_t r = g_xchg[src & 31];
    barrier();
    T out;
    memcpy(&out, &r, sizeof(T));
    return out;
}

}  // namespace emu

template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return emu::exchange(v, src); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned delta) {
    const int lane = (int) threadIdx.x;
    const T o = emu::exchange(v, lane >= (int) delta ? lane - (int) delta : lane);
    return lane >= (int) delta ? o : v;
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu::exchange(v, (int) threadIdx.x ^ m); }
static inline unsigned __ballot_sync(unsigned, int pred) {
    emu::g_xchg[threadIdx.x] = pred ? 1 : 0;
    emu::barrier();
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m |= (unsigned) (emu::g_xchg[i] & 1) << i;
    emu::barrier();
    return m;
}
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::barrier(); }
static inline void __syncthreads() { emu::barrier(); }
static inline void __threadfence_system() {}
static inline void __threadfence() {}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int) v); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline long long clock64() { return 0; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline int atomicAdd(int *p, int v) { int o = *p; *p += v; return o; }
static inline int atomicSub(int *p, int v) { int o = *p; *p -= v; return o; }
static inline int atomicOr(int *p, int v) { int o = *p; *p |= v; return o; }
static inline unsigned atomicOr(unsigned *p, unsigned v) { unsigned o = *p; *p |= v; return o; }
namespace cooperative_groups {
// the set of lanes that happen to be converged: one lane here (each does its own atomic -- same result, not aggregated)
struct lone_thread {
    unsigned size() const { return 1; }
    unsigned thread_rank() const { return 0; }
    template <class T> T shfl(T v, int) const { return v; }
};
inline lone_thread coalesced_threads() { return lone_thread(); }
}  // namespace cooperative_groups
