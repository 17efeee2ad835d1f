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
