// TEST INFRASTRUCTURE: a minimal host emulation of one CUDA thread block (32 ... 256 threads), enough to compile
// i-adhore_b200/csrc/cluster_kernel.cu with g++ and run its kernels on the CPU (GPU-less container).
// The threads are fibers (ucontext) scheduled round-robin inside one OS thread; a warp collective is a
// rendezvous of the 32 lanes of the warp, __syncthreads of the whole block, so divergent code paths between
// collectives behave as on the device as long as the collectives themselves are reached by all their
// participants (true for these kernels).  Arithmetic is the
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
#ifndef EMU_MAX_THREADS
#define EMU_MAX_THREADS 256
#endif
#ifndef EMU_STACK_BYTES
#define EMU_STACK_BYTES (1 << 20)
#endif
constexpr int MAX_THREADS = EMU_MAX_THREADS, MAX_WARPS = MAX_THREADS / 32;
struct Idx { unsigned x, y, z; };
struct Fiber { ucontext_t ctx; std::vector<char> stack; bool done = true; };
extern Fiber fibers[MAX_THREADS];
extern ucontext_t sched_ctx;
extern int cur, n_threads, n_done, n_done_warp[MAX_WARPS];
extern int arrived_block, arrived_warp[MAX_WARPS];
extern unsigned long long block_gen, warp_gen[MAX_WARPS];
extern unsigned block_id, grid_dim;
extern unsigned char xchg[2][MAX_THREADS][16];
extern unsigned long long n_collectives;

inline void yield() { swapcontext(&fibers[cur].ctx, &sched_ctx); }
inline int blocked()
{
    int b = arrived_block;
    for (int w = 0; w < MAX_WARPS; w++) b += arrived_warp[w];
    return b;
}
// Thread 0 leaves a rendezvous only after every other live thread is waiting in its next one.  The device
// code has one pattern that depends on a warp being converged between two barriers: all lanes read a
// value, then thread 0 (the serial sections are `if (tid == 0)`) updates it.  Running thread 0 strictly
// last in every segment reproduces "reads first, then the write".
inline void thread0_last()
{
    if (cur == 0)
        while (blocked() < n_threads - n_done - 1) yield();
}
// __syncthreads: rendezvous of all threads of the block that have not exited
inline void barrier_block()
{
    const unsigned long long my = block_gen;
    if (++arrived_block == n_threads - n_done) {
        arrived_block = 0;
        block_gen++;
        n_collectives++;
    }
    while (block_gen == my) yield();
    thread0_last();
}
// warp collectives: rendezvous of the 32 lanes of the calling thread's warp
inline void barrier_warp()
{
    const int w = cur >> 5;
    const unsigned long long my = warp_gen[w];
    if (++arrived_warp[w] == 32 - n_done_warp[w]) {
        arrived_warp[w] = 0;
        warp_gen[w]++;
        n_collectives++;
    }
    while (warp_gen[w] == my) yield();
    thread0_last();
}
template <typename T>
inline void publish(const T& v)
{
    static_assert(sizeof(T) <= 16, "collective operand too large");
    std::memcpy(xchg[warp_gen[cur >> 5] & 1][cur], &v, sizeof(T));
}
template <typename T>
inline T peek(unsigned long long gen, int lane)   // lane of the calling thread's warp
{
    T v;
    std::memcpy(&v, xchg[gen & 1][(cur & ~31) + lane], sizeof(T));
    return v;
}
inline bool lane_done(int lane) { return fibers[(cur & ~31) + lane].done; }
inline unsigned long long wgen() { return warp_gen[cur >> 5]; }
void run_block(unsigned block, int threads, const std::function<void()>& body);
// a whole grid, block after block (kernels that do not wait for another block: all of K1)
inline void launch(unsigned grid, int threads, const std::function<void()>& body)
{
    grid_dim = grid;
    for (unsigned b = 0; b < grid; b++) run_block(b, threads, body);
    grid_dim = 1;
}
}  // namespace emu

#define threadIdx (emu::Idx{(unsigned)emu::cur, 0u, 0u})
#define blockIdx (emu::Idx{emu::block_id, 0u, 0u})
#define blockDim (emu::Idx{(unsigned)emu::n_threads, 1u, 1u})
#define gridDim (emu::Idx{emu::grid_dim, 1u, 1u})

inline void __syncthreads() { emu::barrier_block(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::barrier_warp(); }
inline void __threadfence() {}
inline long long clock64() { return 0; }

template <typename T>
inline T emu_xchg(T v, const std::function<int(int)>& src)
{
    const unsigned long long gen = emu::wgen();
    emu::publish(v);
    emu::barrier_warp();
    const int s = src(emu::cur & 31);
    return (s >= 0 && s < 32) ? emu::peek<T>(gen, s) : v;
}
template <typename T> inline T __shfl_sync(unsigned, T v, int lane, int = 32) { return emu_xchg<T>(v, [lane](int) { return lane & 31; }); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return emu_xchg<T>(v, [m](int l) { return l ^ m; }); }
template <typename T> inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) { return emu_xchg<T>(v, [d](int l) { return l + (int)d < 32 ? l + (int)d : l; }); }
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) { return emu_xchg<T>(v, [d](int l) { return l - (int)d >= 0 ? l - (int)d : l; }); }
inline unsigned __ballot_sync(unsigned, int pred)
{
    const unsigned long long gen = emu::wgen();
    emu::publish<int>(pred != 0);
    emu::barrier_warp();
    unsigned m = 0;
    for (int l = 0; l < 32; l++)
        if (!emu::lane_done(l) && emu::peek<int>(gen, l)) m |= 1u << l;
    return m;
}
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, !pred) == 0; }
inline int __reduce_min_sync(unsigned, int v)
{
    const unsigned long long gen = emu::wgen();
    emu::publish<int>(v);
    emu::barrier_warp();
    int r = v;
    for (int l = 0; l < 32; l++)
        if (!emu::lane_done(l)) r = std::min(r, emu::peek<int>(gen, l));
    return r;
}
inline int __reduce_max_sync(unsigned, int v)
{
    const unsigned long long gen = emu::wgen();
    emu::publish<int>(v);
    emu::barrier_warp();
    int r = v;
    for (int l = 0; l < 32; l++)
        if (!emu::lane_done(l)) r = std::max(r, emu::peek<int>(gen, l));
    return r;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
// lanes interleave only at collectives, so a plain read-modify-write is atomic here
template <typename T, typename U> inline T atomicAdd(T* p, U v) { const T old = *p; *p = (T)(old + (T)v); return old; }
template <typename T, typename U> inline T atomicMin(T* p, U v) { const T old = *p; if ((T)v < old) *p = (T)v; return old; }
template <typename T, typename U> inline T atomicMax(T* p, U v) { const T old = *p; if ((T)v > old) *p = (T)v; return old; }
template <typename T, typename U, typename V> inline T atomicCAS(T* p, U cmp, V v) { const T old = *p; if (old == (T)cmp) *p = (T)v; return old; }
template <typename T, typename U> inline T atomicOr(T* p, U v) { const T old = *p; *p = (T)(old | (T)v); return old; }

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
