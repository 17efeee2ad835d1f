// K3, first half: per-GHM measuring pass of the multiplicon serialisation (Cluster::setBounds,
// BaseCluster::filterDuplicates, false anchor points of GHMProfile::setMultiplicons, packed sizes).
// A device function so that the clustering kernel runs it as the tail of each GHM (the data is
// still in that SM's L1) -- the separate measure launch and its latency are gone; pack_kernel.cu
// keeps the writer, which needs the global output offsets.
#pragma once
#include <cuda_runtime.h>

#include "common.h"
#include "workspace.h"

namespace pk {

struct PGhm {
    uint32_t* pool;
    ClusterRec* cl;
    MultRec* mult;
    int* out_mults;
    int* oc;       // output clusters in stream order (reuses the `act` array)
    int* scratch;  // per anchor point: false APs before it (reuses the `nxt` array)
    int n_mult, n_oc;
};

__device__ inline void pghm_init(PGhm& p, const GhmTask& task, uint8_t* ws, int pool_factor)
{
    const int D = (int)(task.cnt[0] + task.cnt[1]);
    WsLayout L = ws_layout(D, pool_factor);
    uint8_t* base = ws + task.ws_off;
    p.pool = (uint32_t*)(base + L.pool);
    p.cl = (ClusterRec*)(base + L.cl);
    p.mult = (MultRec*)(base + L.mult);
    p.out_mults = (int*)(base + L.out_mults);
    p.oc = (int*)(base + L.act);
    p.scratch = (int*)(base + L.nxt);
    p.n_mult = task.n_mult;
}

// does profile element (segment n, column x) pair with dataset element yg?  (Gene.cpp:17-30)
// family mode: same family, different gene ID; pairs mode: yg is in the bucket of the element's gene
__device__ __forceinline__ bool seg_match(const DatasetDev& ds, const ProfileDev& pf, int n, int x, int yg, int fy, int gy)
{
    const int idx = n * pf.L + x;
    if (pf.flags[idx] & 2) return false;
    if (ds.pairs) {
        const int f = pf.fam[idx];   // = the element's gene id (upload_profile)
        if (f < 0) return false;
        int lo = ds.fam_ptr[f];
        const int b1 = ds.fam_ptr[f + 1];
        int hi = b1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (ds.fam_members[mid] < yg) lo = mid + 1;
            else hi = mid;
        }
        return lo < b1 && ds.fam_members[lo] == yg;
    }
    return pf.fam[idx] == fy && pf.gid[idx] != gy;
}

// whole CTA (any block size that is a multiple of 32); `p` = one shared PGhm of the CTA, its array
// pointers and n_mult set by the caller (the GHM's global workspace, or the resident kernel's shared memory)
template <bool HL>
__device__ void measure_core(PGhm& p, GhmTask& task, const DatasetDev& ds, const ProfileDev& pf_call)
{
    const ProfileDev& pf = profile_of(pf_call, task.prof);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, PW = blockDim.x >> 5;
    if (tid == 0) {
        int n = 0, sbase = 0;
        for (int m = 0; m < p.n_mult; m++)
            for (int c = p.mult[p.out_mults[m]].head; c >= 0; c = p.cl[c].mnext) {
                p.oc[n++] = c;
                p.cl[c].out_off = sbase;  // temporarily: base of this cluster's scratch segment
                sbase += p.cl[c].n;
            }
        p.n_oc = n;
    }
    __syncthreads();
    const int64_t yoff = ds.list_off[task.y_list];
    for (int i = warp; i < p.n_oc; i += PW) {
        ClusterRec& c = p.cl[p.oc[i]];
        uint32_t* ap = p.pool + c.off;
        int n = c.n;
        int mn = 0x7fffffff, mx = -1;
        bool sorted = true;
        for (int k = lane; k < n; k += 32) {
            const int y = KEY_Y(ap[k]);
            mn = min(mn, y);
            mx = max(mx, y);
            if (HL && k > 0 && ap[k - 1] >= ap[k]) sorted = false;
        }
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, d));
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        }
        int nfalse = 0;
        if (HL) {
            // filterDuplicates: unique (x,y), ordered by x then y.  The array is x-sorted, so a plain
            // sort of the packed keys gives exactly that order; usually it already is.
            if (!__all_sync(0xffffffffu, sorted)) {
                if (lane == 0) {
                    for (int a = 1; a < n; a++) {
                        const uint32_t v = ap[a];
                        int b = a - 1;
                        while (b >= 0 && ap[b] > v) {
                            ap[b + 1] = ap[b];
                            b--;
                        }
                        ap[b + 1] = v;
                    }
                    int w = 1;
                    for (int a = 1; a < n; a++)
                        if (ap[a] != ap[w - 1]) ap[w++] = ap[a];
                    c.n = w;
                }
                __syncwarp();
                n = c.n;
            }
            // false anchor points: every further matching segment of the profile column; the
            // scratch array keeps, per AP, the number of false APs of all earlier APs
            int* scr = p.scratch + c.out_off;
            int run = 0;
            for (int k0 = 0; k0 < n; k0 += 32) {
                const int k = k0 + lane;
                int nf = 0;
                if (k < n) {
                    const int x = KEY_X(ap[k]);
                    const int yg = (int)(yoff + KEY_Y(ap[k]));
                    const int fy = ds.fam[yg], gy = ds.gid[yg];
                    int matches = 0;
                    for (int s = 0; s < pf.S; s++) matches += seg_match(ds, pf, s, x, yg, fy, gy) ? 1 : 0;
                    nf = matches > 1 ? matches - 1 : 0;
                }
                int inc = nf;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, inc, d);
                    if (lane >= d) inc += t;
                }
                if (k < n) scr[k] = run + inc - nf;
                run += __shfl_sync(0xffffffffu, inc, 31);
            }
            nfalse = run;
        }
        if (lane == 0) {
            c.min_y = mn;
            c.max_y = mx;
            c.n_false = nfalse;
            c.has_false = nfalse > 0;
        }
    }
    __syncthreads();
    if (tid == 0) {
        int off = 0, naps = 0;
        for (int m = 0; m < p.n_mult; m++) {
            MultRec& mr = p.mult[p.out_mults[m]];
            mr.out_off = off;
            off += 33;
            for (int c = mr.head; c >= 0; c = p.cl[c].mnext) {
                p.cl[c].cap = p.cl[c].out_off;  // scratch base survives in `cap` (the pool is frozen now)
                p.cl[c].out_off = off;
                off += 102 + 17 * (p.cl[c].n + p.cl[c].n_false);
                naps += p.cl[c].n + p.cl[c].n_false;
            }
        }
        task.out_bytes = off;
        task.n_clusters = p.n_oc;
        task.n_aps = naps;
    }
}

template <bool HL>
__device__ void measure_ghm(PGhm& p, GhmTask& task, uint8_t* ws, const DatasetDev& ds, const ProfileDev& pf,
                            int pool_factor)
{
    if (threadIdx.x == 0) pghm_init(p, task, ws, pool_factor);
    __syncthreads();
    measure_core<HL>(p, task, ds, pf);
}

}  // namespace pk
