// K1: GHM / GHMProfile anchor-point matrix build (family-hash join).
//
// Replaces GHM::buildMatrixFast (GHM.cpp:34-85) and GHMProfile::buildMatrix
// (GHMProfile.cpp:17-59; ListElement::matchingPositions ListElement.cpp:20-30;
// Gene::isPairWith Gene.cpp:17-30).  The reference hashes family *strings* per pair (L2)
// or brute-forces S*L*|Y| string compares (HL); here families are dense ints and a
// dataset-wide CSR (family -> ascending global element indices) is built once, so a
// column's partners inside list Y are one binary search into its family bucket.
//
// Work decomposition: one warp per tile of 32 consecutive columns of one GHM, lane = column.
// Pass 1 counts dots per (tile, orientation) and per GHM; after a per-GHM exclusive scan of
// the tile counts, pass 2 re-walks the buckets and writes the dots as packed keys
// (x<<16|y), already in the (x ascending, y ascending) order of the reference's
// map<int,set<int>> -- no sort needed.  HBM-bound: algorithmic bytes per GHM are
// 5*(|X|+|Y|) element bytes read + 4*D key bytes written.
#include <cuda_runtime.h>

#include <algorithm>

#include "common.h"

namespace {

__device__ __forceinline__ int lower_bound_i32(const int32_t* __restrict__ a, int lo, int hi, int v)
{
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int& total)
{
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    total = __shfl_sync(0xffffffffu, inc, 31);
    return inc - v;
}

// find the GHM owning global tile index t: largest g with tile_prefix[g] <= t
__device__ __forceinline__ int find_ghm(const int64_t* __restrict__ tile_prefix, int n_ghm, int64_t t)
{
    int lo = 0, hi = n_ghm;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (tile_prefix[mid] <= t) lo = mid;
        else hi = mid;
    }
    return lo;
}

// ---------------------------------------------------------------------------------------
// L2: column x of list X against list Y (GHM.cpp:52-84).  Masks are ignored.
template <bool FILL>
__device__ __forceinline__ void l2_column(const DatasetDev& ds, int xg, int x, int64_t yoff, int ylen,
                                          bool identical, int& c0, int& c1, uint32_t* __restrict__ keys,
                                          int64_t o0, int64_t o1)
{
    const int f = ds.fam[xg];
    const int gx = ds.gid[xg];
    const int sx = ds.flags[xg] & 1;
    const int b0 = ds.fam_ptr[f], b1 = ds.fam_ptr[f + 1];
    int m = lower_bound_i32(ds.fam_members, b0, b1, (int)yoff);
    const int yend = (int)(yoff + ylen);
    for (; m < b1; m++) {
        const int yg = ds.fam_members[m];
        if (yg >= yend) break;
        const int y = yg - (int)yoff;
        if (identical && y > x) break;      // GHM.cpp:64 (members ascend, so all later y fail too)
        if (ds.gid[yg] == gx) continue;     // GHM.cpp:68
        if ((ds.flags[yg] & 1) == sx) {
            if (FILL) keys[o1 + c1] = MAKE_KEY(x, y);
            c1++;
        } else {
            if (FILL) keys[o0 + c0] = MAKE_KEY(x, y);
            c0++;
        }
    }
}

// HL: column x of the profile (all segments, merged through the column's groups) against the
// non-masked elements of list Y (GHMProfile.cpp:25-57).  Emits every y once per orientation,
// ascending.
template <bool FILL>
__device__ __forceinline__ void hl_column(const DatasetDev& ds, const ProfileDev& pf, const int16_t* __restrict__ vis, int kprof,
                                          int x, int64_t yoff, int ylen, int& c0, int& c1, uint32_t* __restrict__ keys,
                                          int64_t o0, int64_t o1)
{
    const int g0 = pf.col_ptr[x], g1 = pf.col_ptr[x + 1];
    if (g0 == g1) return;
    const int yend = (int)(yoff + ylen);
    int last = (int)yoff - 1;  // global index of the last candidate handled
    while (true) {
        int best = yend;
        for (int g = g0; g < g1; g++) {
            const int f = pf.grp_fam[g];
            const int b0 = ds.fam_ptr[f], b1 = ds.fam_ptr[f + 1];
            const int m = lower_bound_i32(ds.fam_members, b0, b1, last + 1);
            if (m < b1) {
                const int yg = ds.fam_members[m];
                if (yg < best) best = yg;
            }
        }
        if (best >= yend) break;
        last = best;
        const uint8_t fl = ds.flags[best];
        // masked y elements are invisible (ListElement.cpp:25); in a batch: masked when this profile is searched
        if (vis ? (vis[best] <= kprof) : ((fl & 4) != 0)) continue;
        const int fy = ds.fam[best], gy = ds.gid[best], sy = fl & 1;
        bool e0 = false, e1 = false;
        for (int g = g0; g < g1; g++) {
            if (ds.pairs) {
                // table_type=pairs: a group is one gene; it pairs with `best` iff best is in its bucket
                // (Gene.cpp:24-27) -- buckets overlap, so family equality does not decide it
                const int f = pf.grp_fam[g];
                const int b0 = ds.fam_ptr[f], b1 = ds.fam_ptr[f + 1];
                const int m = lower_bound_i32(ds.fam_members, b0, b1, best);
                if (m >= b1 || ds.fam_members[m] != best) continue;
            } else {
                if (pf.grp_fam[g] != fy) continue;
                const int gg = pf.grp_gid[g];
                if (gg >= 0 && gg == gy) continue;  // Gene.cpp:20: ID != other.ID
            }
            if (pf.grp_strand[g] == sy) e1 = true;
            else e0 = true;
        }
        const int y = best - (int)yoff;
        if (e0) {
            if (FILL) keys[o0 + c0] = MAKE_KEY(x, y);
            c0++;
        }
        if (e1) {
            if (FILL) keys[o1 + c1] = MAKE_KEY(x, y);
            c1++;
        }
    }
}

// ---------------------------------------------------------------------------------------
// L2, second generation: no per-column walk.
//  * the partners of column x inside list Y are the slot range [lo, hi) of its family bucket,
//    read from the (family, list) table (or two binary searches when the table is too large);
//  * for the self comparison the cut `y <= x` (GHM.cpp:64) and the exclusion of the element itself
//    (GHM.cpp:68) together are "slots before the element's own slot";
//  * the per-orientation counts of a range are differences of a strand prefix sum, so the count
//    pass is O(1) per column, and a partner's rank inside its orientation block is a prefix
//    difference too -- every candidate computes its output address independently;
//  * the fill pass expands the candidates of 32 columns over the lanes (load balanced), reads the
//    member records coalesced and writes each column's two key runs contiguously.
// Columns whose gene ID occurs more than once in the dataset (flags bit 3; a shipped dataset has
// none) keep the exact ID test of GHM.cpp:68 on a per-lane slow path.
struct ColRange {
    int lo, hi;   // candidate slots
    int sx;       // strand of the x element
    int slow;     // needs the gene-ID test
    int c0, c1;   // dots per orientation
};

__device__ __forceinline__ ColRange l2_col_range(const DatasetDev& ds, int xg, int y_list, int64_t yoff, int ylen,
                                                 bool identical)
{
    ColRange r;
    const int f = ds.fam[xg];
    const int fl = ds.flags[xg];
    r.sx = fl & 1;
    r.slow = (fl & 8) ? 1 : 0;
    if (ds.fam_list) {
        const int32_t* row = ds.fam_list + (size_t)f * (ds.n_lists + 1) + y_list;
        r.lo = row[0];
        r.hi = row[1];
    } else {
        const int b0 = ds.fam_ptr[f], b1 = ds.fam_ptr[f + 1];
        r.lo = lower_bound_i32(ds.fam_members, b0, b1, (int)yoff);
        r.hi = lower_bound_i32(ds.fam_members, r.lo, b1, (int)(yoff + ylen));
    }
    if (identical) r.hi = ds.self_slot[xg] + (r.slow ? 1 : 0);
    if (!r.slow) {
        const int n = r.hi - r.lo;
        const int np = ds.plus_prefix[r.hi] - ds.plus_prefix[r.lo];
        r.c1 = r.sx ? np : n - np;
        r.c0 = n - r.c1;
    } else {
        const int gx = ds.gid[xg];
        r.c0 = r.c1 = 0;
        for (int m = r.lo; m < r.hi; m++) {
            if (ds.gid[ds.fam_members[m]] == gx) continue;
            if ((int)(ds.mem_y[m] & 1u) == r.sx) r.c1++;
            else r.c0++;
        }
    }
    return r;
}

template <bool FILL>
__global__ void __launch_bounds__(256) l2_build_kernel(DatasetDev ds, GhmTask* __restrict__ tasks, int n_ghm,
                                                       const int64_t* __restrict__ tile_prefix, int64_t n_tiles,
                                                       uint32_t* __restrict__ tile_cnt,
                                                       const uint32_t* __restrict__ tile_off,
                                                       uint32_t* __restrict__ keys)
{
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    // every warp owns a contiguous block of tiles: the owning GHM is searched once, then advanced
    const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t per_warp = (n_tiles + warps_total - 1) / warps_total;
    const int64_t t_begin = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * per_warp;
    const int64_t t_end = (t_begin + per_warp < n_tiles) ? t_begin + per_warp : n_tiles;
    int g = (t_begin < n_tiles) ? find_ghm(tile_prefix, n_ghm, t_begin) : 0;
    for (int64_t t = t_begin; t < t_end; t++) {
        while (t >= tile_prefix[g + 1]) g++;
        const int x_list = tasks[g].x_list, y_list = tasks[g].y_list;
        const int x0 = (int)(t - tile_prefix[g]) * 32;
        const int64_t yoff = ds.list_off[y_list];
        const int ylen = (int)(ds.list_off[y_list + 1] - yoff);
        const int64_t xoff = ds.list_off[x_list];
        const int nx = (int)(ds.list_off[x_list + 1] - xoff);
        const int x = x0 + lane;
        ColRange r;
        r.lo = r.hi = r.sx = r.slow = r.c0 = r.c1 = 0;
        if (x < nx) r = l2_col_range(ds, (int)(xoff + x), y_list, yoff, ylen, x_list == y_list);
        int t0, t1;
        const int e0 = warp_excl_scan(r.c0, lane, t0);
        const int e1 = warp_excl_scan(r.c1, lane, t1);
        if (!FILL) {
            if (lane == 0) {
                tile_cnt[2 * t] = (uint32_t)t0;
                tile_cnt[2 * t + 1] = (uint32_t)t1;
                if (t0) atomicAdd(&tasks[g].cnt[0], (uint32_t)t0);
                if (t1) atomicAdd(&tasks[g].cnt[1], (uint32_t)t1);
            }
            continue;
        }
        if (t0 + t1 == 0) continue;
        // offsets inside this GHM's key block stay 32-bit (a GHM holds < 2^28 dots)
        uint32_t* __restrict__ kb = keys + tasks[g].dot_off;
        const int base0 = (int)tile_off[2 * t];
        const int base1 = (int)(tasks[g].cnt[0] + tile_off[2 * t + 1]);
        // slow columns: the lane walks its own range
        if (r.slow) {
            const int gx = ds.gid[(int)(xoff + x)];
            int o0 = base0 + e0, o1 = base1 + e1;
            for (int m = r.lo; m < r.hi; m++) {
                if (ds.gid[ds.fam_members[m]] == gx) continue;
                const uint32_t my = ds.mem_y[m];
                const uint32_t key = MAKE_KEY(x, my >> 1);
                if ((int)(my & 1u) == r.sx) kb[o1++] = key;
                else kb[o0++] = key;
            }
        }
        // fast columns: groups of W lanes take one column each (W = the power of two below the tile's
        // mean candidates per column), so a group reads its column's member records coalesced and
        // writes the column's two key runs in order; no per-candidate search.  W = 1 (sparse tiles)
        // degenerates to one lane per column, W = 32 to the whole warp per column.
        const int len = r.slow ? 0 : r.hi - r.lo;
        int total;
        warp_excl_scan(len, lane, total);
        if (total == 0) continue;
        int W = 1;
        while (W < 32 && 2 * W <= (total >> 5)) W <<= 1;
        const int R = 32 / W;
        const int sub = lane & (W - 1), grp = lane / W;
        const int pp_lo = ds.plus_prefix[r.lo];
        for (int cb = 0; cb < 32; cb += R) {
            const int c = cb + grp;
            const int c_len = __shfl_sync(full, len, c);
            const int maxlen = __reduce_max_sync(full, c_len);
            if (maxlen == 0) continue;
            const int c_lo = __shfl_sync(full, r.lo, c);
            const int c_pp = __shfl_sync(full, pp_lo, c);
            const int c_sx = __shfl_sync(full, r.sx, c);
            const int c_b0 = base0 + __shfl_sync(full, e0, c);
            const int c_b1 = base1 + __shfl_sync(full, e1, c);
            const uint32_t xkey = (uint32_t)(x0 + c) << 16;
            // two candidates per lane and trip: both record loads are in flight before either is used
            for (int k = sub; k < c_len; k += 2 * W) {
                const int k2 = k + W;
                const bool two = k2 < c_len;
                const int m = c_lo + k;
                const uint32_t my = ds.mem_y[m];
                const int pp = ds.plus_prefix[m];
                const uint32_t my2 = two ? ds.mem_y[m + W] : 0u;
                const int pp2 = two ? ds.plus_prefix[m + W] : 0;
                {
                    const int plus_before = pp - c_pp;
                    const int sy = (int)(my & 1u);
                    const int rank = sy ? plus_before : k - plus_before;
                    kb[((sy == c_sx) ? c_b1 : c_b0) + rank] = xkey | (my >> 1);
                }
                if (two) {
                    const int plus_before = pp2 - c_pp;
                    const int sy = (int)(my2 & 1u);
                    const int rank = sy ? plus_before : k2 - plus_before;
                    kb[((sy == c_sx) ? c_b1 : c_b0) + rank] = xkey | (my2 >> 1);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// L2, third generation, for batches of many sparse GHMs (config 4: 125,250 list pairs, ~420 dots each,
// 2,500 columns of which five in six are empty in any one GHM).  The tile kernels above visit every
// (GHM, column) cell -- 313 M of them for 53 M dots.  Here the work is the dots themselves:
//  * a column (element e of list X, family f) meets every member of f's bucket once: a member in list Y is
//    a dot of the GHM (X, Y) if the call holds that pair (pair_idx, an n_lists x n_lists table of task
//    indices, -1 = not in this call / this rank's share); the self comparison keeps the slots before the
//    element's own (y < x, GHM.cpp:64-68), columns with a repeated gene ID keep the exact ID test;
//  * pass 1 counts per (GHM, orientation) with one reduction per dot, the host turns the counts into
//    offsets (it needs them anyway, for the clustering plan), pass 2 walks again and appends each key
//    to its GHM's orientation block through an atomic cursor -- in arbitrary order;
//  * pass 3 sorts every block ascending in shared memory: keys are (x << 16) | y and distinct, so the sorted
//    block IS the reference's map<int, set<int>> order, independent of the order of arrival.
// Member records are read coalesced (groups of W lanes per column, W adapted to the mean bucket length).
template <bool FILL>
__global__ void __launch_bounds__(256) l2_expand_kernel(DatasetDev ds, int64_t n_elems, const int32_t* __restrict__ pair_idx,
                                                        const uint8_t* __restrict__ row_used,
                                                        uint32_t* __restrict__ counter /* [2n]: counts (pass 1) or cursors (pass 2) */,
                                                        const int64_t* __restrict__ dot_off, const uint32_t* __restrict__ cnt,
                                                        uint32_t* __restrict__ keys)
{
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t n_tiles = (n_elems + 31) >> 5;
    const int nl = ds.n_lists;
    for (int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < n_tiles; t += warps_total) {
        const int64_t e = t * 32 + lane;
        int len = 0, b0 = 0, X = 0, self = 0, meta = 0, gx = 0, x = 0;
        if (e < n_elems) {
            X = ds.elem_list[e];
            if (row_used[X]) {   // (lists that are no GHM's x list in this call have nothing to do)
                const int f = ds.fam[e];
                const int fl = ds.flags[e];
                b0 = ds.fam_ptr[f];
                len = ds.fam_ptr[f + 1] - b0;
                const int slow = (fl & 8) ? 1 : 0;
                self = ds.self_slot[e] + slow;          // the self comparison keeps slots < self
                meta = (fl & 1) | (slow << 1);
                if (slow) gx = ds.gid[e];
                x = (int)(e - ds.list_off[X]);
            }
        }
        int total;
        warp_excl_scan(len, lane, total);
        if (total == 0) continue;
        int W = 1;
        while (W < 32 && 2 * W <= (total >> 5)) W <<= 1;
        const int R = 32 / W;
        const int sub = lane & (W - 1), grp = lane / W;
        for (int cb = 0; cb < 32; cb += R) {
            const int c = cb + grp;
            const int c_len = __shfl_sync(full, len, c);
            if (__reduce_max_sync(full, c_len) == 0) continue;
            const int c_b0 = __shfl_sync(full, b0, c);
            const int c_X = __shfl_sync(full, X, c);
            const int c_self = __shfl_sync(full, self, c);
            const int c_meta = __shfl_sync(full, meta, c);
            const int c_gx = __shfl_sync(full, gx, c);
            const uint32_t xkey = (uint32_t)__shfl_sync(full, x, c) << 16;
            const int32_t* __restrict__ row = pair_idx + (size_t)c_X * nl;
            // two members per lane and trip: both records are in flight before either is used
            auto emit = [&](int m, int Y, int g, uint32_t my) {
                if (g < 0) return;
                if (Y == c_X && m >= c_self) return;
                if ((c_meta & 2) && ds.gid[ds.fam_members[m]] == c_gx) return;   // GHM.cpp:68
                const int o = ((int)(my & 1u) == (c_meta & 1)) ? 1 : 0;
                if (!FILL) atomicAdd(&counter[2 * g + o], 1u);
                else {
                    const uint32_t pos = atomicAdd(&counter[2 * g + o], 1u);
                    keys[dot_off[g] + (o ? cnt[2 * g] : 0u) + pos] = xkey | (my >> 1);
                }
            };
            for (int k = sub; k < c_len; k += 2 * W) {
                const int m = c_b0 + k;
                const bool two = k + W < c_len;
                const int Y = ds.mem_list[m];
                const int Y2 = two ? ds.mem_list[m + W] : 0;
                const uint32_t my = ds.mem_y[m];
                const uint32_t my2 = two ? ds.mem_y[m + W] : 0u;
                const int g = row[Y];
                const int g2 = two ? row[Y2] : -1;
                emit(m, Y, g, my);
                emit(m + W, Y2, g2, my2);
            }
        }
    }
}

// The same for profile GHMs (family mode): the work item is a column group of a profile -- one (family, strand) of
// one aligned column, gid = its single gene or -1 (common.h) -- which meets every member of its family's bucket once:
// a non-masked member in list Y, not the group's own gene (Gene.cpp:20), is the dot (column, y) of the GHM
// (profile, Y), same orientation iff the strands agree.  A column's groups of one family differ in strand, so they
// feed different orientation blocks of a GHM: no y is produced twice (the dedupe of GHMProfile.cpp:52-55).
// Pairs mode keeps the tile kernel (buckets of different genes overlap there).
template <bool FILL>
__global__ void __launch_bounds__(256) hl_expand_kernel(DatasetDev ds, const HlGroup* __restrict__ items, int n_items,
                                                        const int32_t* __restrict__ task_idx, const int16_t* __restrict__ vis,
                                                        uint32_t* __restrict__ counter, const int64_t* __restrict__ dot_off,
                                                        const uint32_t* __restrict__ cnt, uint32_t* __restrict__ keys)
{
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    const int warps_total = gridDim.x * (blockDim.x >> 5);
    const int n_tiles = (n_items + 31) >> 5;
    const int nl = ds.n_lists;
    for (int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < n_tiles; t += warps_total) {
        const int it = t * 32 + lane;
        HlGroup G;
        G.fam = G.gid = G.prof = 0;
        G.x_strand = 0;
        int len = 0, b0 = 0;
        if (it < n_items) {
            G = items[it];
            b0 = ds.fam_ptr[G.fam];
            len = ds.fam_ptr[G.fam + 1] - b0;
        }
        int total;
        warp_excl_scan(len, lane, total);
        if (total == 0) continue;
        int W = 1;
        while (W < 32 && 2 * W <= (total >> 5)) W <<= 1;
        const int R = 32 / W;
        const int sub = lane & (W - 1), grp = lane / W;
        for (int cb = 0; cb < 32; cb += R) {
            const int c = cb + grp;
            const int c_len = __shfl_sync(full, len, c);
            if (__reduce_max_sync(full, c_len) == 0) continue;
            const int c_b0 = __shfl_sync(full, b0, c);
            const int c_gid = __shfl_sync(full, G.gid, c);
            const int c_prof = __shfl_sync(full, G.prof, c);
            const uint32_t c_xs = __shfl_sync(full, G.x_strand, c);
            const uint32_t xkey = (c_xs & 0xffffu) << 16;
            const int c_strand = (int)(c_xs >> 31);
            const int32_t* __restrict__ row = task_idx + (size_t)c_prof * nl;
            for (int k = sub; k < c_len; k += W) {
                const int m = c_b0 + k;
                const int g = row[ds.mem_list[m]];
                if (g < 0) continue;
                const int e = ds.fam_members[m];
                if (vis ? (vis[e] <= c_prof) : ((ds.flags[e] & 4) != 0)) continue;   // ListElement.cpp:25
                if (c_gid >= 0 && ds.gid[e] == c_gid) continue;                       // Gene.cpp:20
                const uint32_t my = ds.mem_y[m];
                const int o = ((int)(my & 1u) == c_strand) ? 1 : 0;
                if (!FILL) atomicAdd(&counter[2 * g + o], 1u);
                else {
                    const uint32_t pos = atomicAdd(&counter[2 * g + o], 1u);
                    keys[dot_off[g] + (o ? cnt[2 * g] : 0u) + pos] = xkey | (my >> 1);
                }
            }
        }
    }
}

// Ascending sort of every (GHM, orientation) key block (the engine takes this path only when the largest block
// fits SORT_MAX keys).  Blocks of up to 1,024 keys -- all but a handful on the sparse batches this path serves --
// are sorted by ONE WARP IN REGISTERS: the block is padded with 0xffffffff to P = 32 R keys, lane l holds the keys at
// positions l, l + 32, ... (coalesced loads and stores); a compare-exchange at distance j >= 32 pairs two registers
// of the same lane, at distance j < 32 the same register of lanes l and l ^ j (one shuffle).  No shared memory, no
// barrier; the eight warps of a CTA work on different blocks.  Larger blocks: one CTA, bitonic network in shared memory.
#define SORT_MAX 8192
#define SORT_WARP_MAX 1024
template <int R>
__device__ __forceinline__ void warp_sort_regs(uint32_t* __restrict__ kb, int n, int lane)
{
    uint32_t v[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int p = lane + 32 * r;
        v[r] = p < n ? kb[p] : 0xffffffffu;
    }
#pragma unroll
    for (int k = 2; k <= 32 * R; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int jr = j >> 5;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    if (r & jr) continue;
                    // positions p = lane + 32 r (bit j clear) and p + j; ascending where (p & k) == 0
                    const bool up = (((lane + 32 * r) & k) == 0);
                    const uint32_t a = v[r], b = v[r | jr];
                    const uint32_t lo = min(a, b), hi = max(a, b);
                    v[r] = up ? lo : hi;
                    v[r | jr] = up ? hi : lo;
                }
            } else {
                const bool lower = ((lane & j) == 0);
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const uint32_t w = __shfl_xor_sync(0xffffffffu, v[r], j);
                    const bool up = (((lane + 32 * r) & k) == 0);
                    v[r] = (lower == up) ? min(v[r], w) : max(v[r], w);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int p = lane + 32 * r;
        if (p < n) kb[p] = v[r];
    }
}

// (two instantiations, launched one after the other: blocks of up to 256 keys -- 8 registers per lane, many warps
// resident -- and blocks of 257 ... 1,024 keys)
template <int RMAX>
__global__ void __launch_bounds__(256) seg_sort_warp_kernel(int n_seg, const int64_t* __restrict__ dot_off,
                                                            const uint32_t* __restrict__ cnt, uint32_t* __restrict__ keys)
{
    const int lane = threadIdx.x & 31;
    const int warps_total = gridDim.x * (blockDim.x >> 5);
    for (int sg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); sg < n_seg; sg += warps_total) {
        const int n = (int)cnt[sg];
        if (RMAX == 8 ? (n < 2 || n > 256) : (n <= 256 || n > SORT_WARP_MAX)) continue;
        uint32_t* __restrict__ kb = keys + dot_off[sg >> 1] + ((sg & 1) ? cnt[sg & ~1] : 0u);
        if (RMAX == 8) {
            if (n <= 32) warp_sort_regs<1>(kb, n, lane);
            else if (n <= 64) warp_sort_regs<2>(kb, n, lane);
            else if (n <= 128) warp_sort_regs<4>(kb, n, lane);
            else warp_sort_regs<8>(kb, n, lane);
        } else {
            if (n <= 512) warp_sort_regs<16>(kb, n, lane);
            else warp_sort_regs<32>(kb, n, lane);
        }
    }
}

__global__ void __launch_bounds__(256) seg_sort_cta_kernel(int n_seg, const int64_t* __restrict__ dot_off,
                                                           const uint32_t* __restrict__ cnt, uint32_t* __restrict__ keys)
{
    __shared__ uint32_t s[SORT_MAX];
    const int tid = threadIdx.x;
    for (int sg = blockIdx.x; sg < n_seg; sg += gridDim.x) {
        const int n = (int)cnt[sg];
        if (n <= SORT_WARP_MAX) continue;
        uint32_t* __restrict__ kb = keys + dot_off[sg >> 1] + ((sg & 1) ? cnt[sg & ~1] : 0u);
        int P = 2048;
        while (P < n) P <<= 1;
        __syncthreads();   // (the previous block's readers are done with s)
        for (int i = tid; i < P; i += 256) s[i] = i < n ? kb[i] : 0xffffffffu;
        __syncthreads();
        for (int k = 2; k <= P; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < (P >> 1); i += 256) {
                    // the i-th compare-exchange of this step: positions a < b = a + j
                    const int a = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                    const int b = a | j;
                    const uint32_t va = s[a], vb = s[b];
                    const bool up = ((a & k) == 0);
                    if ((va > vb) == up) {
                        s[a] = vb;
                        s[b] = va;
                    }
                }
                __syncthreads();
            }
        for (int i = tid; i < n; i += 256) kb[i] = s[i];
    }
}

// ---------------------------------------------------------------------------------------
// Dataset tables on the device (family mode; iadh_create): the family CSR and everything K1 reads per member slot are
// built from the flat element arrays by kernels -- the join side of DataSet::getGeneFamilies / GHM::buildMatrixFast
// (the y-side map<family, vector<element>> of GHM.cpp:37-50, built once for the whole dataset instead of per pair).
//   count    family histogram + gene-ID histogram (an ID seen twice switches K1 to its exact ID test, flags bit 3)
//   scan     exclusive prefix sum -> fam_ptr
//   scatter  element -> a slot of its family's bucket (atomic cursor, arbitrary order inside the bucket)
//   sort     every bucket ascending (warp-register bitonic; shared-memory / global bitonic for large families):
//            members ascending by global element index, as the tile kernels' binary searches and the (family, list)
//            table expect
//   records  per slot: (coordinate << 1 | strand), list; per element: own slot; strand prefix sum
// exclusive scan of n int32 (out has n + 1 entries, out[n] = total): block-local scan, scan of the block sums, add
__global__ void __launch_bounds__(1024) scan_block_kernel(const int32_t* __restrict__ in, int64_t n, int32_t* __restrict__ out,
                                                          int32_t* __restrict__ block_sum)
{
    __shared__ int32_t warp_sum[32];
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int32_t v = i < n ? in[i] : 0;
    int total;
    int32_t excl = warp_excl_scan(v, lane, total);
    if (lane == 31) warp_sum[w] = total;
    __syncthreads();
    if (w == 0) {
        int t2;
        const int32_t x = warp_sum[lane];
        const int32_t e = warp_excl_scan(x, lane, t2);
        warp_sum[lane] = e;
        if (lane == 31) block_sum[blockIdx.x] = t2;
    }
    __syncthreads();
    if (i < n) out[i] = warp_sum[w] + excl;
}
__global__ void __launch_bounds__(1024) scan_sums_kernel(int32_t* block_sum, int n_blocks)
{
    __shared__ int32_t warp_sum[32];
    __shared__ int32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < n_blocks; base += 1024) {
        const int i = base + threadIdx.x;
        const int32_t v = i < n_blocks ? block_sum[i] : 0;
        int total;
        const int32_t excl = warp_excl_scan(v, lane, total);
        if (lane == 31) warp_sum[w] = total;
        __syncthreads();
        if (w == 0) {
            int t2;
            warp_sum[lane] = warp_excl_scan(warp_sum[lane], lane, t2);
        }
        __syncthreads();
        const int32_t e = carry + warp_sum[w] + excl;
        if (i < n_blocks) block_sum[i] = e;
        __syncthreads();
        if (threadIdx.x == 1023) carry = e + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sum[n_blocks] = carry;   // grand total
}
__global__ void __launch_bounds__(256) scan_add_kernel(int32_t* __restrict__ out, int64_t n, const int32_t* __restrict__ block_sum, int n_blocks)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += block_sum[i >> 10];
    if (i == n) out[n] = block_sum[n_blocks];
}

__global__ void __launch_bounds__(256) tab_count_kernel(const int32_t* __restrict__ fam, const int32_t* __restrict__ gid, int64_t n,
                                                        int32_t* fam_cnt, int32_t* gid_cnt)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    atomicAdd(&fam_cnt[fam[i]], 1);
    const int g = gid[i];
    if (g >= 0) atomicAdd(&gid_cnt[g], 1);
}
__global__ void __launch_bounds__(256) tab_scatter_kernel(const int32_t* __restrict__ fam, const int64_t* __restrict__ list_off, int n_lists,
                                                          const int32_t* __restrict__ gid, const int32_t* __restrict__ gid_cnt, int64_t n,
                                                          const int32_t* __restrict__ fam_ptr, int32_t* cursor, int32_t* __restrict__ members,
                                                          int32_t* __restrict__ elem_list, uint8_t* __restrict__ flags)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int f = fam[i];
    members[fam_ptr[f] + atomicAdd(&cursor[f], 1)] = (int32_t)i;
    int lo = 0, hi = n_lists;   // the list of element i: largest l with list_off[l] <= i
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (list_off[mid] <= i) lo = mid;
        else hi = mid;
    }
    elem_list[i] = lo;
    const int g = gid[i];
    if (g < 0 || gid_cnt[g] > 1) flags[i] |= 8;   // the gene ID occurs more than once: exact ID test in K1
}
// segments of a CSR: bucket f = keys[ptr[f] .. ptr[f + 1])
template <int RMAX>
__global__ void __launch_bounds__(256) csr_sort_warp_kernel(int n_seg, const int32_t* __restrict__ ptr, uint32_t* __restrict__ keys)
{
    const int lane = threadIdx.x & 31;
    const int warps_total = gridDim.x * (blockDim.x >> 5);
    for (int sg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); sg < n_seg; sg += warps_total) {
        const int b = ptr[sg], n = ptr[sg + 1] - b;
        if (RMAX == 8 ? (n < 2 || n > 256) : (n <= 256 || n > SORT_WARP_MAX)) continue;
        uint32_t* __restrict__ kb = keys + b;
        if (RMAX == 8) {
            if (n <= 32) warp_sort_regs<1>(kb, n, lane);
            else if (n <= 64) warp_sort_regs<2>(kb, n, lane);
            else if (n <= 128) warp_sort_regs<4>(kb, n, lane);
            else warp_sort_regs<8>(kb, n, lane);
        } else {
            if (n <= 512) warp_sort_regs<16>(kb, n, lane);
            else warp_sort_regs<32>(kb, n, lane);
        }
    }
}
// buckets of more than 1,024 members: one CTA each; up to SORT_MAX keys in shared memory, beyond that on global memory
// with the all-ascending form of the network (every merge starts with a "flip" stage, i against k - 1 - i, then plain
// half-cleaners): every compare-exchange moves the smaller key down, so the +infinity padding above position n never
// moves and stays virtual -- pairs that reach beyond n are skipped
__global__ void __launch_bounds__(256) csr_sort_cta_kernel(int n_seg, const int32_t* __restrict__ ptr, uint32_t* __restrict__ keys)
{
    __shared__ uint32_t s[SORT_MAX];
    const int tid = threadIdx.x;
    for (int sg = blockIdx.x; sg < n_seg; sg += gridDim.x) {
        const int b = ptr[sg], n = ptr[sg + 1] - b;
        if (n <= SORT_WARP_MAX) continue;
        uint32_t* __restrict__ kb = keys + b;
        int64_t P = 2048;
        while (P < n) P <<= 1;
        __syncthreads();
        if (P <= SORT_MAX) {
            for (int i = tid; i < (int)P; i += 256) s[i] = i < n ? kb[i] : 0xffffffffu;
            __syncthreads();
            for (int k = 2; k <= (int)P; k <<= 1)
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int i = tid; i < ((int)P >> 1); i += 256) {
                        const int a = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                        const int c = a | j;
                        const uint32_t va = s[a], vb = s[c];
                        const bool up = ((a & k) == 0);
                        if ((va > vb) == up) {
                            s[a] = vb;
                            s[c] = va;
                        }
                    }
                    __syncthreads();
                }
            for (int i = tid; i < n; i += 256) kb[i] = s[i];
            continue;
        }
        for (int64_t k = 2; k <= P; k <<= 1) {
            const int64_t h = k >> 1;
            for (int64_t i = tid; i < (P >> 1); i += 256) {   // flip
                const int64_t blk = i / h, off = i - blk * h;
                const int64_t a = blk * k + off, c = blk * k + (k - 1 - off);
                if (c < n) {
                    const uint32_t va = kb[a], vb = kb[c];
                    if (va > vb) {
                        kb[a] = vb;
                        kb[c] = va;
                    }
                }
            }
            __syncthreads();
            for (int64_t j = k >> 2; j > 0; j >>= 1) {
                for (int64_t i = tid; i < (P >> 1); i += 256) {
                    const int64_t a = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                    const int64_t c = a | j;
                    if (c < n) {
                        const uint32_t va = kb[a], vb = kb[c];
                        if (va > vb) {
                            kb[a] = vb;
                            kb[c] = va;
                        }
                    }
                }
                __syncthreads();
            }
        }
    }
}
__global__ void __launch_bounds__(256) tab_records_kernel(const int32_t* __restrict__ members, int64_t n_mem, const int32_t* __restrict__ elem_list,
                                                          const int64_t* __restrict__ list_off, const uint8_t* __restrict__ flags,
                                                          uint32_t* __restrict__ mem_y, int32_t* __restrict__ mem_list,
                                                          int32_t* __restrict__ self_slot, int32_t* __restrict__ strand)
{
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_mem) return;
    const int e = members[m];
    const int l = elem_list[e];
    const int st = flags[e] & 1;
    mem_y[m] = ((uint32_t)(e - list_off[l]) << 1) | (uint32_t)st;
    mem_list[m] = l;
    self_slot[e] = (int32_t)m;
    strand[m] = st;
}

// task records <- compact per-GHM arrays (family-expansion path: counts were taken outside the records)
__global__ void scatter_cnt_kernel(GhmTask* __restrict__ tasks, int n, const uint32_t* __restrict__ cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        tasks[i].cnt[0] = cnt[2 * i];
        tasks[i].cnt[1] = cnt[2 * i + 1];
    }
}

template <bool HL, bool FILL>
__global__ void __launch_bounds__(256) build_kernel(DatasetDev ds, ProfileDev pf, GhmTask* __restrict__ tasks,
                                                    int n_ghm, const int64_t* __restrict__ tile_prefix,
                                                    int64_t n_tiles, uint32_t* __restrict__ tile_cnt,
                                                    const uint32_t* __restrict__ tile_off,
                                                    uint32_t* __restrict__ keys)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < n_tiles; t += warps_total) {
        const int g = find_ghm(tile_prefix, n_ghm, t);
        const GhmTask task = tasks[g];
        const ProfileDev& P = profile_of(pf, task.prof);
        const int x0 = (int)(t - tile_prefix[g]) * 32;
        const int64_t yoff = ds.list_off[task.y_list];
        const int ylen = (int)(ds.list_off[task.y_list + 1] - yoff);
        int nx;
        int64_t xoff = 0;
        if (HL) nx = P.L;
        else {
            xoff = ds.list_off[task.x_list];
            nx = (int)(ds.list_off[task.x_list + 1] - xoff);
        }
        const int x = x0 + lane;
        int c0 = 0, c1 = 0;
        if (!FILL) {
            if (x < nx) {
                if (HL) hl_column<false>(ds, P, pf.vis, task.prof, x, yoff, ylen, c0, c1, nullptr, 0, 0);
                else l2_column<false>(ds, (int)(xoff + x), x, yoff, ylen, task.x_list == task.y_list, c0, c1, nullptr, 0, 0);
            }
            int t0, t1;
            warp_excl_scan(c0, lane, t0);
            warp_excl_scan(c1, lane, t1);
            // per-lane counts are kept (2 x uint16 would overflow for huge families): the fill
            // pass recomputes them; only tile totals are stored.
            if (lane == 0) {
                tile_cnt[2 * t] = (uint32_t)t0;
                tile_cnt[2 * t + 1] = (uint32_t)t1;
                if (t0) atomicAdd(&tasks[g].cnt[0], (uint32_t)t0);
                if (t1) atomicAdd(&tasks[g].cnt[1], (uint32_t)t1);
            }
        } else {
            // pass 2: lane offsets = tile offset + exclusive scan of the recomputed lane counts
            if (x < nx) {
                if (HL) hl_column<false>(ds, P, pf.vis, task.prof, x, yoff, ylen, c0, c1, nullptr, 0, 0);
                else l2_column<false>(ds, (int)(xoff + x), x, yoff, ylen, task.x_list == task.y_list, c0, c1, nullptr, 0, 0);
            }
            int t0, t1;
            const int e0 = warp_excl_scan(c0, lane, t0);
            const int e1 = warp_excl_scan(c1, lane, t1);
            if (t0 + t1 == 0) continue;
            const int64_t base0 = task.dot_off + tile_off[2 * t];
            const int64_t base1 = task.dot_off + task.cnt[0] + tile_off[2 * t + 1];
            c0 = c1 = 0;
            if (x < nx) {
                if (HL) hl_column<true>(ds, P, pf.vis, task.prof, x, yoff, ylen, c0, c1, keys, base0 + e0, base1 + e1);
                else l2_column<true>(ds, (int)(xoff + x), x, yoff, ylen, task.x_list == task.y_list, c0, c1, keys, base0 + e0, base1 + e1);
            }
        }
    }
}

// per-GHM exclusive scan of the tile counts (one warp per GHM)
__global__ void __launch_bounds__(256) tile_scan_kernel(const int64_t* __restrict__ tile_prefix, int n_ghm,
                                                        const uint32_t* __restrict__ tile_cnt,
                                                        uint32_t* __restrict__ tile_off)
{
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= n_ghm) return;
    const int64_t t0 = tile_prefix[g], t1 = tile_prefix[g + 1];
    int run0 = 0, run1 = 0;
    for (int64_t t = t0; t < t1; t += 32) {
        const int64_t tt = t + lane;
        int a = 0, b = 0;
        if (tt < t1) {
            a = (int)tile_cnt[2 * tt];
            b = (int)tile_cnt[2 * tt + 1];
        }
        int ta, tb;
        const int ea = warp_excl_scan(a, lane, ta);
        const int eb = warp_excl_scan(b, lane, tb);
        if (tt < t1) {
            tile_off[2 * tt] = (uint32_t)(run0 + ea);
            tile_off[2 * tt + 1] = (uint32_t)(run1 + eb);
        }
        run0 += ta;
        run1 += tb;
    }
}

// (family, list) -> first slot of the family at or after the list (built once per dataset)
__global__ void __launch_bounds__(128) fam_list_kernel(const int32_t* __restrict__ fam_ptr,
                                                       const int32_t* __restrict__ members,
                                                       const int64_t* __restrict__ list_off, int n_fam, int n_lists,
                                                       int64_t total, int32_t* __restrict__ table)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_fam) return;
    int m = fam_ptr[f];
    const int end = fam_ptr[f + 1];
    int32_t* row = table + (size_t)f * (n_lists + 1);
    for (int l = 0; l <= n_lists; l++) {
        const int64_t lim = (l < n_lists) ? list_off[l] : total;
        while (m < end && members[m] < lim) m++;
        row[l] = m;
    }
}

// GeneList::mask (GeneList.cpp:283-287)
__global__ void mask_kernel(uint8_t* flags, int64_t begin, int64_t end, int set)
{
    int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= end) flags[i] = set ? (flags[i] | 4) : (flags[i] & ~4);
}

}  // namespace

// ---------------------------------------------------------------------------------------
// host launchers

#ifndef IADH_HOST_EMU   // (tests/emu compiles the kernels above for the host and launches them itself)
static inline int grid_for(int64_t warps_needed, int sm_count)
{
    int64_t blocks = (warps_needed + 7) / 8;
    int64_t cap = (int64_t)sm_count * 8;  // 8 CTAs x 256 threads = 2048 threads per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

void launch_build_count(bool hl, const DatasetDev& ds, const ProfileDev& pf, GhmTask* tasks, int n_ghm,
                        const int64_t* tile_prefix, int64_t n_tiles, uint32_t* tile_cnt, int sm_count,
                        cudaStream_t st)
{
    if (n_tiles == 0) return;
    int grid = grid_for(n_tiles, sm_count);
    if (hl) build_kernel<true, false><<<grid, 256, 0, st>>>(ds, pf, tasks, n_ghm, tile_prefix, n_tiles, tile_cnt, nullptr, nullptr);
    else l2_build_kernel<false><<<grid, 256, 0, st>>>(ds, tasks, n_ghm, tile_prefix, n_tiles, tile_cnt, nullptr, nullptr);
}

void launch_tile_scan(const int64_t* tile_prefix, int n_ghm, const uint32_t* tile_cnt, uint32_t* tile_off,
                      cudaStream_t st)
{
    if (n_ghm == 0) return;
    tile_scan_kernel<<<(n_ghm + 7) / 8, 256, 0, st>>>(tile_prefix, n_ghm, tile_cnt, tile_off);
}

void launch_build_fill(bool hl, const DatasetDev& ds, const ProfileDev& pf, GhmTask* tasks, int n_ghm,
                       const int64_t* tile_prefix, int64_t n_tiles, const uint32_t* tile_off, uint32_t* keys,
                       int sm_count, cudaStream_t st)
{
    if (n_tiles == 0) return;
    int grid = grid_for(n_tiles, sm_count);
    if (hl) build_kernel<true, true><<<grid, 256, 0, st>>>(ds, pf, tasks, n_ghm, tile_prefix, n_tiles, nullptr, tile_off, keys);
    else l2_build_kernel<true><<<grid, 256, 0, st>>>(ds, tasks, n_ghm, tile_prefix, n_tiles, nullptr, tile_off, keys);
}

static void launch_seg_sort(int n_seg, const int64_t* dot_off, const uint32_t* cnt, uint32_t* keys, int sm_count, cudaStream_t st)
{
    seg_sort_warp_kernel<8><<<std::min((n_seg + 7) / 8, sm_count * 16), 256, 0, st>>>(n_seg, dot_off, cnt, keys);
    seg_sort_warp_kernel<32><<<std::min((n_seg + 7) / 8, sm_count * 8), 256, 0, st>>>(n_seg, dot_off, cnt, keys);
    seg_sort_cta_kernel<<<std::min(n_seg, sm_count * 4), 256, 0, st>>>(n_seg, dot_off, cnt, keys);
}

// family-expansion path (many sparse GHMs): count / fill / sort
void launch_expand_count(const DatasetDev& ds, int64_t n_elems, const int32_t* pair_idx, const uint8_t* row_used, uint32_t* cnt,
                         int sm_count, cudaStream_t st)
{
    if (n_elems == 0) return;
    l2_expand_kernel<false><<<grid_for((n_elems + 31) / 32, sm_count), 256, 0, st>>>(ds, n_elems, pair_idx, row_used, cnt, nullptr, nullptr, nullptr);
}

void launch_expand_fill(const DatasetDev& ds, int64_t n_elems, const int32_t* pair_idx, const uint8_t* row_used, uint32_t* cursor,
                        const int64_t* dot_off, const uint32_t* cnt, uint32_t* keys, GhmTask* tasks, int n_ghm, int sm_count,
                        cudaStream_t st)
{
    if (n_ghm == 0) return;
    scatter_cnt_kernel<<<(n_ghm + 255) / 256, 256, 0, st>>>(tasks, n_ghm, cnt);
    if (n_elems == 0) return;
    l2_expand_kernel<true><<<grid_for((n_elems + 31) / 32, sm_count), 256, 0, st>>>(ds, n_elems, pair_idx, row_used, cursor, dot_off, cnt, keys);
    const int n_seg = 2 * n_ghm;
    launch_seg_sort(n_seg, dot_off, cnt, keys, sm_count, st);
}
void launch_hl_expand_count(const DatasetDev& ds, const HlGroup* items, int n_items, const int32_t* task_idx, const int16_t* vis,
                            uint32_t* cnt, int sm_count, cudaStream_t st)
{
    if (n_items == 0) return;
    hl_expand_kernel<false><<<grid_for((n_items + 31) / 32, sm_count), 256, 0, st>>>(ds, items, n_items, task_idx, vis, cnt, nullptr, nullptr, nullptr);
}

void launch_hl_expand_fill(const DatasetDev& ds, const HlGroup* items, int n_items, const int32_t* task_idx, const int16_t* vis,
                           uint32_t* cursor, const int64_t* dot_off, const uint32_t* cnt, uint32_t* keys, GhmTask* tasks, int n_ghm,
                           int sm_count, cudaStream_t st)
{
    if (n_ghm == 0) return;
    scatter_cnt_kernel<<<(n_ghm + 255) / 256, 256, 0, st>>>(tasks, n_ghm, cnt);
    if (n_items == 0) return;
    hl_expand_kernel<true><<<grid_for((n_items + 31) / 32, sm_count), 256, 0, st>>>(ds, items, n_items, task_idx, vis, cursor, dot_off, cnt, keys);
    const int n_seg = 2 * n_ghm;
    launch_seg_sort(n_seg, dot_off, cnt, keys, sm_count, st);
}
int expand_sort_max(void) { return SORT_MAX; }

// exclusive scan, out[n] = total; block_sum holds n / 1024 + 2 entries
void launch_scan_i32(const int32_t* in, int64_t n, int32_t* out, int32_t* block_sum, cudaStream_t st)
{
    const int n_blocks = (int)((n + 1023) / 1024);
    if (n_blocks > 0) scan_block_kernel<<<n_blocks, 1024, 0, st>>>(in, n, out, block_sum);
    scan_sums_kernel<<<1, 1024, 0, st>>>(block_sum, n_blocks);
    scan_add_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(out, n, block_sum, n_blocks);
}

// Dataset tables of iadh_create on the device (family mode).  In: fam / gid / flags (strand, mask bits) / list_off on
// the device; zeroed fam_cnt[n_fam + 1], gid_cnt[n_gid + 1], cursor[n_fam + 1].  Out: fam_ptr, members, elem_list,
// mem_y, mem_list, self_slot, plus_prefix, flags bit 3.  tmp = n ints, block_sum = n / 1024 + 2 ints.
void launch_tables(const int32_t* fam, const int32_t* gid, uint8_t* flags, const int64_t* list_off, int n_lists, int64_t n, int n_fam,
                   int32_t* fam_cnt, int32_t* gid_cnt, int32_t* cursor, int32_t* fam_ptr, int32_t* members, int32_t* elem_list,
                   uint32_t* mem_y, int32_t* mem_list, int32_t* self_slot, int32_t* plus_prefix, int32_t* tmp, int32_t* block_sum,
                   int sm_count, cudaStream_t st)
{
    if (n == 0) {
        cudaMemsetAsync(fam_ptr, 0, 4 * (size_t)(n_fam + 1), st);
        cudaMemsetAsync(plus_prefix, 0, 4, st);
        return;
    }
    const unsigned g256 = (unsigned)((n + 255) / 256);
    tab_count_kernel<<<g256, 256, 0, st>>>(fam, gid, n, fam_cnt, gid_cnt);
    launch_scan_i32(fam_cnt, n_fam, fam_ptr, block_sum, st);
    tab_scatter_kernel<<<g256, 256, 0, st>>>(fam, list_off, n_lists, gid, gid_cnt, n, fam_ptr, cursor, members, elem_list, flags);
    csr_sort_warp_kernel<8><<<std::min((n_fam + 7) / 8, sm_count * 16), 256, 0, st>>>(n_fam, fam_ptr, (uint32_t*)members);
    csr_sort_warp_kernel<32><<<std::min((n_fam + 7) / 8, sm_count * 8), 256, 0, st>>>(n_fam, fam_ptr, (uint32_t*)members);
    csr_sort_cta_kernel<<<std::min(n_fam, sm_count * 4), 256, 0, st>>>(n_fam, fam_ptr, (uint32_t*)members);
    tab_records_kernel<<<g256, 256, 0, st>>>(members, n, elem_list, list_off, flags, mem_y, mem_list, self_slot, tmp);
    launch_scan_i32(tmp, n, plus_prefix, block_sum, st);
}

void launch_mask(uint8_t* flags, int64_t begin, int64_t end, int set, cudaStream_t st)
{
    int64_t n = end - begin + 1;
    if (n <= 0) return;
    mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(flags, begin, end, set);
}

void launch_fam_list(const int32_t* fam_ptr, const int32_t* members, const int64_t* list_off, int n_fam, int n_lists,
                     int64_t total, int32_t* table, cudaStream_t st)
{
    if (n_fam == 0) return;
    fam_list_kernel<<<(n_fam + 127) / 128, 128, 0, st>>>(fam_ptr, members, list_off, n_fam, n_lists, total, table);
}
#endif  // IADH_HOST_EMU
