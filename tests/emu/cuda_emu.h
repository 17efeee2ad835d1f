// TEST INFRASTRUCTURE: a minimal host emulation of one CUDA thread block of ONE warp, enough to compile
// i-adhore_b200/csrc/cluster_kernel.cu with g++ and run its kernels on the CPU (GPU-less container).
// The 32 lanes are fibers (ucontext) scheduled round-robin inside one OS thread; every warp collective and
// barrier is a rendezvous of all 32, so divergent code paths between collectives behave as on the device as
// long as the collectives themselves are reached by all lanes (true for these kernels).  Arithmetic is the
// host's IEEE double arithmetic, compiled with -ffp-contract=off (the device code is built --fmad=false).
// Not a performance model and not a race detector: lanes only interleave at collectives.
#pragma once
#include <cuda_runtime.h>   // vector types (int2, double2, ...) and the empty host definitions of __device__ etc.
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <type_traits>
#include <vector>

#undef __shared__
#define __shared__ static
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __noinline__
#define __noinline__ __attribute__((noinline))
#ifndef __forceinline__
#define __forceinline__ inline
#endif

namespace emu {
constexpr int LANES = 32;
struct Idx { unsigned x, y, z; };
struct Fiber { ucontext_t ctx; std::vector<char> stack; bool done = true; };
extern Fiber fibers[LANES];
extern ucontext_t sched_ctx;
extern int cur, n_done, arrived;
extern unsigned long long generation;
extern unsigned block_id;
extern unsigned char xchg[2][LANES][16];
extern unsigned long long n_collectives;

inline void yield() { swapcontext(&fibers[cur].ctx, &sched_ctx); }
// rendezvous of all lanes that have not exited
inline void barrier()
{
    const unsigned long long my = generation;
    if (++arrived == LANES - n_done) {
        arrived = 0;
        generation++;
        n_collectives++;
    }
    while (generation == my) yield();
    // Lane 0 leaves a rendezvous only after every other live lane has reached the next one.  The device
    // code has one pattern that depends on the warp being converged between two barriers: all lanes
    // read a value, then lane 0 (the serial sections are `if (tid == 0)`) updates it.  Running lane 0
    // strictly last in every segment reproduces "reads first, then the write".
    if (cur == 0)
        while (arrived < LANES - n_done - 1) yield();
}
template <typename T>
inline void publish(const T& v)
{
    static_assert(sizeof(T) <= 16, "collective operand too large");
    std::memcpy(xchg[generation & 1][cur], &v, sizeof(T));
}
template <typename T>
inline T peek(unsigned long long gen, int lane)
{
    T v;
    std::memcpy(&v, xchg[gen & 1][lane], sizeof(T));
    return v;
}
void run_block(unsigned block, const std::function<void()>& body);
}  // namespace emu

#define threadIdx (emu::Idx{(unsigned)emu::cur, 0u, 0u})
#define blockIdx (emu::Idx{emu::block_id, 0u, 0u})
#define blockDim (emu::Idx{32u, 1u, 1u})
#define gridDim (emu::Idx{1u, 1u, 1u})

inline void __syncthreads() { emu::barrier(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::barrier(); }
inline void __threadfence() {}
inline long long clock64() { return 0; }

template <typename T>
inline T emu_xchg(T v, const std::function<int(int)>& src)
{
    const unsigned long long gen = emu::generation;
    emu::publish(v);
    emu::barrier();
    const int s = src(emu::cur);
    return (s >= 0 && s < emu::LANES) ? emu::peek<T>(gen, s) : v;
}
template <typename T> inline T __shfl_sync(unsigned, T v, int lane, int = 32) { return emu_xchg<T>(v, [lane](int) { return lane & 31; }); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return emu_xchg<T>(v, [m](int l) { return l ^ m; }); }
template <typename T> inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) { return emu_xchg<T>(v, [d](int l) { return l + (int)d < 32 ? l + (int)d : l; }); }
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) { return emu_xchg<T>(v, [d](int l) { return l - (int)d >= 0 ? l - (int)d : l; }); }
inline unsigned __ballot_sync(unsigned, int pred)
{
    const unsigned long long gen = emu::generation;
    emu::publish<int>(pred != 0);
    emu::barrier();
    unsigned m = 0;
    for (int l = 0; l < emu::LANES; l++)
        if (!emu::fibers[l].done && emu::peek<int>(gen, l)) m |= 1u << l;
    return m;
}
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, !pred) == 0; }
inline int __reduce_min_sync(unsigned, int v)
{
    const unsigned long long gen = emu::generation;
    emu::publish<int>(v);
    emu::barrier();
    int r = v;
    for (int l = 0; l < emu::LANES; l++) r = std::min(r, emu::peek<int>(gen, l));
    return r;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
// lanes interleave only at collectives, so a plain read-modify-write is atomic here
template <typename T, typename U> inline T atomicAdd(T* p, U v) { const T old = *p; *p = (T)(old + (T)v); return old; }

// CUDA's global min / max overload set
template <class A, class B> inline typename std::common_type<A, B>::type min(A a, B b)
{
    typedef typename std::common_type<A, B>::type T;
    return (T)b < (T)a ? (T)b : (T)a;
}
template <class A, class B> inline typename std::common_type<A, B>::type max(A a, B b)
{
    typedef typename std::common_type<A, B>::type T;
    return (T)a < (T)b ? (T)b : (T)a;
}
