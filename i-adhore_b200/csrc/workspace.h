// Per-GHM clustering workspace layout (bytes), shared by host sizing and the kernels.
// D = dots of both orientations after the build.  Every array is sized from D:
//   key/alive : the mutable matrix per orientation (capacity D each: Bonferroni re-insertion
//               can move a dot into the other orientation's matrix, GHM.cpp:456-459)
//   pool      : anchor-point arrays of all base clusters (bump allocated, pool_factor * D)
//   clusters  : at most D/3 seeds (every cluster starts with >= 3 distinct dots)
#pragma once
#include <stdint.h>

#include "common.h"

struct WsLayout {
    size_t key0, key1, alive0, alive1, nxt, pool, pflag, cl, bc0, bc1, act, act_next, changed, mult, out_mults, rb_d, rb_j, rank, pos, redo, xr, dscr, tmp;
    size_t total;
    int pool_cap, cl_cap, tmp_cap;
};

#ifdef __CUDACC__
__host__ __device__
#endif
static inline WsLayout ws_layout(int D, int pool_factor)
{
    WsLayout L;
    size_t o = 0;
    auto take = [&o](size_t bytes) {
        size_t r = o;
        o += (bytes + 15) & ~(size_t)15;
        return r;
    };
    const size_t d = (size_t)D;
    L.cl_cap = D / 3 + 4;   // two regions of D_o / 3 + 2 slots (one per orientation)
    L.pool_cap = pool_factor * D + 64;
    int tc = 1;
    while (tc < D) tc <<= 1;
    L.tmp_cap = tc;
    L.key0 = take(4 * d);
    L.key1 = take(4 * d);
    L.alive0 = take(d);
    L.alive1 = take(d);
    L.nxt = take(4 * d);      // seed successor pointers; region of orientation 1 starts at D0
    L.pool = take(4 * (size_t)L.pool_cap);
    L.pflag = take((size_t)L.pool_cap);
    L.cl = take(sizeof(ClusterRec) * (size_t)L.cl_cap);
    L.bc0 = take(4 * (size_t)L.cl_cap);
    L.bc1 = take(4 * (size_t)L.cl_cap);
    L.act = take(4 * (size_t)L.cl_cap);
    L.act_next = take(4 * (size_t)L.cl_cap);
    L.changed = take((size_t)L.cl_cap);
    L.mult = take(sizeof(MultRec) * (size_t)L.cl_cap);
    L.out_mults = take(4 * (size_t)L.cl_cap);
    L.rb_d = take(8 * (size_t)L.cl_cap);
    L.rb_j = take(4 * (size_t)L.cl_cap);
    L.rank = take(4 * (size_t)L.cl_cap);
    L.pos = take(4 * (size_t)L.cl_cap);
    L.redo = take(4 * (size_t)L.cl_cap);
    L.xr = take(8 * (size_t)L.cl_cap);   // per active cluster: x-range widened by the gap (enrich / join prefilter)
    L.dscr = take(16 * d + 32);   // per-AP terms of the ordered statistics sums
    L.tmp = take(4 * (size_t)L.tmp_cap + d + 32);  // staging: keys + one flag byte per entry
    L.total = (o + 255) & ~(size_t)255;
    return L;
}
