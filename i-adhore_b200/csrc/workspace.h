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
    // (callers reject GHMs whose pool would not fit an int: ws_pool_fits below)
    L.pool_cap = (int)((long long)pool_factor * D + 64);
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

// ---- shared-memory resident layout (resident_kernel, cluster_kernel.cu) ----------------------------
// The same arrays as WsLayout, sized for a GHM whose whole clustering state lives in the shared memory
// of ONE CTA from the first seed to the packed sizes: no workspace traffic to HBM between passes.
// Capacities are budgets, not bounds -- a GHM that outgrows one (cluster slots, anchor-point pool, matrix
// re-insertion) reports a status and is re-run by the host on the global-memory path; results never
// depend on the budget.  Differences from WsLayout:
//   * cluster slots are recycled after merges, so cl_cap bounds the LIVE clusters (config 3 / 4: at most
//     D / 10 live against D / 4 ever created);
//   * the pool is 2 D entries (tight block capacities + compaction) instead of 8 D;
//   * dscr is a fixed chunk (the ordered sums are accumulated chunk by chunk);
//   * the two orientations are clustered one after the other by the same CTA: act / xr / ... are shared.
// pool_cap, ClusterRec::off and ::cap are 32-bit: a (GHM, pool factor) pair is usable only below that
static inline bool ws_pool_fits(long long D, int pool_factor) { return (long long)pool_factor * D + 64 <= 0x7fffffffll; }

struct RsLayout {
    uint32_t key0, key1, alive0, alive1, nxt, tmp, tmpf, pool, pflag, cl, bc0, bc1, act, act_next, changed, mult, out_mults, rb_d, rb_j,
        rank, pos, redo, xr, dscr;
    uint32_t total;
    int cap0, cap1, tmp_cap, pool_cap, cl_cap0, cl_cap1, dscr_cap;
};

#ifdef __CUDACC__
__host__ __device__
#endif
static inline RsLayout rs_layout(int D0, int D1, int nt, int squeeze = 0)
{
    RsLayout L;
    uint32_t o = 0;
    auto take = [&o](uint32_t bytes) {
        const uint32_t r = o;
        o += (bytes + 15u) & ~15u;
        return r;
    };
    const int D = D0 + D1;
    L.cap0 = D0 + 32;   // matrix capacity: the dots + room for re-inserted anchor points of the other orientation
    L.cap1 = D1 + 32;
    L.tmp_cap = (L.cap0 > L.cap1 ? L.cap0 : L.cap1);
    L.pool_cap = 2 * D + 64;
    int c0 = D0 / 8 + 8, c1 = D1 / 8 + 8;
    if (squeeze) {   // test knob (IADH_RES_SQUEEZE): budgets that real GHMs outgrow, to exercise the fallback
        L.pool_cap = D / 2 + 16;
        c0 = c1 = 3;
        L.cap0 = D0;
        L.cap1 = D1;
    }
    L.cl_cap0 = c0 < D0 / 3 + 2 ? c0 : D0 / 3 + 2;   // (never more slots than the global table has)
    L.cl_cap1 = c1 < D1 / 3 + 2 ? c1 : D1 / 3 + 2;
    const uint32_t cc = (uint32_t)(L.cl_cap0 + L.cl_cap1);
    L.dscr_cap = nt > 64 ? nt : 64;
    L.cl = take((uint32_t)sizeof(ClusterRec) * cc);
    L.mult = take((uint32_t)sizeof(MultRec) * cc);
    L.rb_d = take(8 * cc);
    L.xr = take(8 * cc);
    L.dscr = take(16 * (uint32_t)L.dscr_cap);
    L.key0 = take(4 * (uint32_t)L.cap0);
    L.key1 = take(4 * (uint32_t)L.cap1);
    L.nxt = take(4 * (uint32_t)(L.cap0 + L.cap1));   // successor pointers persist across the gap sizes (seed_pass)
    L.tmp = take(4 * (uint32_t)L.tmp_cap);
    L.pool = take(4 * (uint32_t)L.pool_cap);
    L.bc0 = take(4 * (uint32_t)L.cl_cap0);
    L.bc1 = take(4 * (uint32_t)L.cl_cap1);
    L.act = take(4 * cc);
    L.act_next = take(4 * cc);
    L.out_mults = take(4 * cc);
    L.rb_j = take(4 * cc);
    L.rank = take(4 * cc);
    L.pos = take(4 * cc);
    L.redo = take(4 * cc);
    L.alive0 = take((uint32_t)L.cap0);
    L.alive1 = take((uint32_t)L.cap1);
    L.tmpf = take((uint32_t)L.tmp_cap);
    L.pflag = take((uint32_t)L.pool_cap);
    L.changed = take(cc);
    L.total = o;
    return L;
}
