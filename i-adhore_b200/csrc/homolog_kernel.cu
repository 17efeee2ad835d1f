// Homolog-count filter of a profile search on the device.
//
// Replaces the last step of GHMProfile::setMultiplicons (GHMProfile.cpp:129-132): a multiplicon found by searching a
// profile in a gene list survives only if Multiplicon::createHomologs(useFamily, minHomologs) (Multiplicon.cpp:352-382)
// finds at least minHomologs elements "with a homolog" in EVERY segment of the multiplicon -- the S profile segments cut
// to [begin_x, end_x] and the list cut to [begin_y, end_y] (extractXObject / extractYObject, Multiplicon.cpp:331-350).
// An element has a homolog if
//   (a) pruneLinks (:384-421): it is an end of a Link inherited from the profile's multiplicon whose OTHER end is still
//       inside the cut (both ends are profile elements, i.e. both columns must lie in [begin_x, end_x]), or
//   (b) addLinks (:423-494): an x-segment element and a y-segment element are homologous -- same family and not the
//       same list element (:484-485) / in pairs mode the y gene is in the x gene's pair set (:437) -- which marks both.
// Here: the inherited Links of a profile are resolved to columns once per search (a hash table (segment, numID) -> column
// over the profile elements, one probe per Link end) and turned into a CSR "element -> columns of its Link partners"
// (count, scan, fill); the filter kernel then evaluates (a) by a range
// test over an element's few partners and (b) by intersecting the element's family bucket (the dataset's CSR, members
// ascending) with [begin_y, end_y], which marks the y side in a shared-memory bitmap at the same time.  A multiplicon
// costs S * (end_x - begin_x + 1) bucket look-ups instead of the reference's S * |x| * |y| string compares.  Dropped
// multiplicons are unlinked from the GHM's output list and the packed offsets / sizes of the measuring pass
// (pack_measure.cuh) are recomputed, so the writer (pack_kernel.cu) and the host never see them.
#include <cuda_runtime.h>

#include <algorithm>

#include "common.h"
#include "workspace.h"

#include "homolog.h"
#include "pack_measure.cuh"

namespace {

constexpr int HT = 128;                                  // threads of a filter CTA
constexpr int YW = (IADH_MAX_COORD + 31) / 32 + 1;      // bitmap words for the longest possible y range

// ---- link tables -----------------------------------------------------------------------------------------------
// Links arrive as the caller's array (include/iadhore_b200.h: iadh_link_t = LinkIn), the Links of profile k at
// [link_ptr[k], link_ptr[k+1]); the elements of the profiles this rank searches are numbered consecutively, profile k
// from elem_base[k] (segment-major inside a profile, like ProfileDev::num_id).

// the profile an index of a CSR-like prefix array belongs to: largest k with ptr[k] <= i (empty ranges are skipped)
template <typename T>
__device__ __forceinline__ int owner_of(const T* __restrict__ ptr, int n, int64_t i)
{
    int lo = 0, hi = n;   // invariant: ptr[lo] <= i < ptr[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if ((int64_t)ptr[mid] <= i) lo = mid;
        else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ uint32_t hash_key(unsigned long long k)
{
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return (uint32_t)k;
}
// (first element of the segment, numID) identifies a profile element: numIDs are unique per list element
// (DataSet.cpp:76-81); should a segment ever hold an id twice, the later column wins, as in the map pruneLinks
// builds (Multiplicon.cpp:397)
__device__ __forceinline__ unsigned long long elem_key(int seg_first, int id) { return ((unsigned long long)(uint32_t)seg_first << 32) | (uint32_t)id; }

// open-addressing table (key -> column) over all non-gap elements; keys start as ~0, vals as -1
__global__ void __launch_bounds__(256) link_hash_kernel(const int32_t* __restrict__ elem_base, int n_prof, int64_t n_elems, ProfileDev pf_call,
                                                        unsigned long long* keys, int32_t* vals, uint32_t mask)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_elems) return;
    const int k = owner_of(elem_base, n_prof, e);
    const ProfileDev& pf = profile_of(pf_call, k);
    const int local = (int)(e - elem_base[k]);
    if (pf.flags[local] & 2) return;   // gap
    const int col = local % pf.L;
    const unsigned long long key = elem_key(elem_base[k] + (local - col), pf.num_id[local]);
    uint32_t slot = hash_key(key) & mask;
    while (true) {
        const unsigned long long old = atomicCAS(&keys[slot], ~0ull, key);
        if (old == ~0ull || old == key) {
            atomicMax(&vals[slot], col);
            return;
        }
        slot = (slot + 1) & mask;
    }
}

// per Link end: the column of its gene in its segment, or -1 (gene not in the profile / segment index out of range)
__global__ void __launch_bounds__(256) link_resolve_kernel(const LinkIn* __restrict__ links, const int64_t* __restrict__ link_ptr,
                                                           const int32_t* __restrict__ elem_base, int n_prof, ProfileDev pf_call,
                                                           const unsigned long long* __restrict__ keys, const int32_t* __restrict__ vals,
                                                           uint32_t mask, int32_t* __restrict__ cols)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= 2 * link_ptr[n_prof]) return;
    const LinkIn lk = links[w >> 1];
    const int k = owner_of(link_ptr, n_prof, w >> 1);
    const ProfileDev& pf = profile_of(pf_call, k);
    const int seg = (w & 1) ? lk.seg_y : lk.seg_x, id = (w & 1) ? lk.gene_y : lk.gene_x;
    int found = -1;
    if (seg >= 0 && seg < pf.S) {
        const unsigned long long key = elem_key(elem_base[k] + seg * pf.L, id);
        uint32_t slot = hash_key(key) & mask;
        while (true) {
            const unsigned long long cur = keys[slot];
            if (cur == key) {
                found = vals[slot];
                break;
            }
            if (cur == ~0ull) break;
            slot = (slot + 1) & mask;
        }
    }
    cols[w] = found;
}

// per profile element: number of inherited Links it is an end of (Links with an unresolved end never survive)
__global__ void __launch_bounds__(256) link_count_kernel(const LinkIn* __restrict__ links, const int64_t* __restrict__ link_ptr,
                                                         const int32_t* __restrict__ elem_base, int n_prof, ProfileDev pf_call,
                                                         const int32_t* __restrict__ cols, int32_t* cnt)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= link_ptr[n_prof]) return;
    const int cx = cols[2 * i], cy = cols[2 * i + 1];
    if (cx < 0 || cy < 0) return;
    const LinkIn lk = links[i];
    const int k = owner_of(link_ptr, n_prof, i);
    const int L = profile_of(pf_call, k).L;
    atomicAdd(&cnt[(size_t)elem_base[k] + (size_t)lk.seg_x * L + cx], 1);
    atomicAdd(&cnt[(size_t)elem_base[k] + (size_t)lk.seg_y * L + cy], 1);
}

// CSR fill: every element gets the columns of its partners (order inside an element's run is arbitrary: the filter
// only asks whether ANY partner lies in a column range)
__global__ void __launch_bounds__(256) link_fill_kernel(const LinkIn* __restrict__ links, const int64_t* __restrict__ link_ptr,
                                                        const int32_t* __restrict__ elem_base, int n_prof, ProfileDev pf_call,
                                                        const int32_t* __restrict__ cols, const int32_t* __restrict__ lk_ptr,
                                                        int32_t* cursor, int32_t* __restrict__ lk_col)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= link_ptr[n_prof]) return;
    const int cx = cols[2 * i], cy = cols[2 * i + 1];
    if (cx < 0 || cy < 0) return;
    const LinkIn lk = links[i];
    const int k = owner_of(link_ptr, n_prof, i);
    const int L = profile_of(pf_call, k).L;
    const size_t ex = (size_t)elem_base[k] + (size_t)lk.seg_x * L + cx, ey = (size_t)elem_base[k] + (size_t)lk.seg_y * L + cy;
    lk_col[lk_ptr[ex] + atomicAdd(&cursor[ex], 1)] = cy;
    lk_col[lk_ptr[ey] + atomicAdd(&cursor[ey], 1)] = cx;
}

// One CTA per profile x list GHM, after the clustering kernels and the measuring pass, before the outcome read-back.
__global__ void __launch_bounds__(HT) homolog_filter_kernel(GhmTask* tasks, int n_tasks, uint8_t* ws, DatasetDev ds, int n_fam,
                                                            ProfileDev pf_call, const HomologDev* __restrict__ hdev, int pool_factor,
                                                            int min_homologs)
{
    __shared__ pk::PGhm p;
    __shared__ int* keep;            // per output multiplicon: survives (the join's `rank` array, idle by now)
    __shared__ uint32_t ybits[YW];   // y elements of [begin_y, end_y] with a homolog
    __shared__ int sh_b[4];          // begin_x, end_x, begin_y, end_y of the multiplicon (Multiplicon.cpp:135-193)
    __shared__ int sh_cnt;
    const int tid = threadIdx.x;
    for (int t = blockIdx.x; t < n_tasks; t += gridDim.x) {
        GhmTask& task = tasks[t];
        __syncthreads();   // (the previous GHM's readers of the shared state are done)
        if (task.status != 0 || task.n_mult <= 0) continue;
        if (tid == 0) {
            pk::pghm_init(p, task, ws, pool_factor);
            const WsLayout WL = ws_layout((int)(task.cnt[0] + task.cnt[1]), pool_factor);
            keep = (int*)(ws + task.ws_off + WL.rank);
        }
        __syncthreads();
        const ProfileDev& pf = profile_of(pf_call, task.prof);
        const HomologDev hd = hdev[task.prof];
        const int64_t yoff = ds.list_off[task.y_list];
        const int S = pf.S, L = pf.L;
        for (int m = 0; m < p.n_mult; m++) {
            if (tid == 0) {
                const MultRec& mr = p.mult[p.out_mults[m]];
                int bx = 0, ex = 0, by = 0, ey = 0;
                bool first = true;
                for (int c = mr.head; c >= 0; c = p.cl[c].mnext) {
                    const ClusterRec& cr = p.cl[c];
                    const int lx = KEY_X(p.pool[cr.off]), hx = KEY_X(p.pool[cr.off + cr.n - 1]);
                    if (first) {
                        bx = lx; ex = hx; by = cr.min_y; ey = cr.max_y;
                        first = false;
                    } else {
                        bx = min(bx, lx); ex = max(ex, hx); by = min(by, cr.min_y); ey = max(ey, cr.max_y);
                    }
                }
                sh_b[0] = bx; sh_b[1] = ex; sh_b[2] = by; sh_b[3] = ey;
            }
            __syncthreads();
            const int bx = sh_b[0], ex = sh_b[1], by = sh_b[2], ey = sh_b[3];
            const int y_words = (ey - by + 32) >> 5;
            for (int w = tid; w < y_words; w += HT) ybits[w] = 0;
            bool failed = false;
            // x-segments, one after the other (Multiplicon.cpp:361-372); (b) marks the y side on the way
            for (int s = 0; s < S; s++) {
                if (tid == 0) sh_cnt = 0;
                __syncthreads();   // (also orders the bitmap reset before the first marks)
                int local = 0;
                for (int col = bx + tid; col <= ex; col += HT) {
                    const size_t idx = (size_t)s * L + col;
                    if (pf.flags[idx] & 2) continue;   // gap
                    bool has = false;
                    for (int k = hd.lk_ptr[idx]; k < hd.lk_ptr[idx + 1]; k++) {   // (a)
                        const int q = hd.lk_col[k];
                        if (q >= bx && q <= ex) {
                            has = true;
                            break;
                        }
                    }
                    const int f = pf.fam[idx];   // family; pairs mode: the element's gene id (bucket = its partners)
                    if (f >= 0 && f < n_fam) {   // (b)
                        const int b1 = ds.fam_ptr[f + 1];
                        int lo = ds.fam_ptr[f], hi = b1;
                        const int64_t y_first = yoff + by, y_last = yoff + ey;
                        while (lo < hi) {
                            const int mid = (lo + hi) >> 1;
                            if (ds.fam_members[mid] < y_first) lo = mid + 1;
                            else hi = mid;
                        }
                        const int my_id = pf.num_id[idx];
                        for (; lo < b1; lo++) {
                            const int64_t yg = ds.fam_members[lo];
                            if (yg > y_last) break;
                            if (!ds.pairs && ds.num_id[yg] == my_id) continue;   // the same list element (Multiplicon.cpp:484-485)
                            has = true;
                            const int bit = (int)(yg - y_first);
                            atomicOr(&ybits[bit >> 5], 1u << (bit & 31));
                        }
                    }
                    local += has ? 1 : 0;
                }
                if (local) atomicAdd(&sh_cnt, local);
                __syncthreads();
                const int count = sh_cnt;
                __syncthreads();
                if (count < min_homologs) {
                    failed = true;
                    break;
                }
            }
            if (!failed) {   // the y-segment (Multiplicon.cpp:374-381)
                if (tid == 0) sh_cnt = 0;
                __syncthreads();
                int local = 0;
                for (int w = tid; w < y_words; w += HT) local += __popc(ybits[w]);
                if (local) atomicAdd(&sh_cnt, local);
                __syncthreads();
                failed = sh_cnt < min_homologs;
            }
            if (tid == 0) keep[m] = failed ? 0 : 1;
            __syncthreads();
        }
        // unlink the dropped multiplicons; offsets and sizes as the measuring pass lays them out (pack_measure.cuh)
        if (tid == 0) {
            int w = 0, noc = 0, off = 0, naps = 0;
            for (int m = 0; m < p.n_mult; m++) {
                if (!keep[m]) continue;
                const int mi = p.out_mults[m];
                p.out_mults[w++] = mi;
                MultRec& mr = p.mult[mi];
                mr.out_off = off;
                off += 33;
                for (int c = mr.head; c >= 0; c = p.cl[c].mnext) {
                    p.oc[noc++] = c;
                    p.cl[c].out_off = off;
                    off += 102 + 17 * (p.cl[c].n + p.cl[c].n_false);
                    naps += p.cl[c].n + p.cl[c].n_false;
                }
            }
            task.n_mult = w;
            task.out_bytes = off;
            task.n_clusters = noc;
            task.n_aps = naps;
        }
    }
}

}  // namespace

#ifndef IADH_HOST_EMU
void launch_scan_i32(const int32_t* in, int64_t n, int32_t* out, int32_t* block_sum, cudaStream_t st);

// links -> per-element partner columns.  keys / vals: hash table of `mask + 1` slots (a power of two >= 2 * n_elems),
// cols: 2 * n_links ints, cnt: n_elems ints (zeroed here, reused as the fill cursor), lk_ptr: n_elems + 1,
// lk_col: 2 * n_links, block_sum: n_elems / 1024 + 2
void launch_link_tables(const LinkIn* links, const int64_t* link_ptr, int64_t n_links, const int32_t* elem_base, int n_prof, int64_t n_elems,
                        const ProfileDev& pf, unsigned long long* keys, int32_t* vals, uint32_t mask, int32_t* cols, int32_t* cnt,
                        int32_t* lk_ptr, int32_t* lk_col, int32_t* block_sum, cudaStream_t st)
{
    cudaMemsetAsync(cnt, 0, 4 * (size_t)n_elems + 4, st);
    cudaMemsetAsync(keys, 0xff, 8 * ((size_t)mask + 1), st);
    cudaMemsetAsync(vals, 0xff, 4 * ((size_t)mask + 1), st);
    const unsigned gl = (unsigned)((n_links + 255) / 256), ge = (unsigned)((n_elems + 255) / 256);
    if (n_elems > 0) link_hash_kernel<<<ge, 256, 0, st>>>(elem_base, n_prof, n_elems, pf, keys, vals, mask);
    if (n_links > 0) {
        link_resolve_kernel<<<(unsigned)((2 * n_links + 255) / 256), 256, 0, st>>>(links, link_ptr, elem_base, n_prof, pf, keys, vals, mask, cols);
        link_count_kernel<<<gl, 256, 0, st>>>(links, link_ptr, elem_base, n_prof, pf, cols, cnt);
    }
    launch_scan_i32(cnt, n_elems, lk_ptr, block_sum, st);
    if (n_links > 0) {
        cudaMemsetAsync(cnt, 0, 4 * (size_t)n_elems + 4, st);
        link_fill_kernel<<<gl, 256, 0, st>>>(links, link_ptr, elem_base, n_prof, pf, cols, lk_ptr, cnt, lk_col);
    }
}

void launch_homolog_filter(GhmTask* tasks, int n_tasks, uint8_t* ws, const DatasetDev& ds, int n_fam, const ProfileDev& pf,
                           const HomologDev* hdev, int pool_factor, int min_homologs, int sm_count, cudaStream_t st)
{
    if (n_tasks <= 0) return;
    homolog_filter_kernel<<<std::min(n_tasks, sm_count * 8), HT, 0, st>>>(tasks, n_tasks, ws, ds, n_fam, pf, hdev, pool_factor, min_homologs);
}
#endif  // IADH_HOST_EMU
