// K3: multiplicon assembly and serialisation into the reference's wire format.
//
// Replaces GHM::setMultiplicons (GHM.cpp:671-701) and the device-computable part of
// GHMProfile::setMultiplicons (GHMProfile.cpp:68-127): Cluster::setBounds (Cluster.h:27-33),
// anchor-point gene ids, BaseCluster::filterDuplicates (BaseCluster.cpp:63-75) and the
// "false" anchor points of further matching profile segments; then
// Multiplicon::pack / BaseCluster::pack / AnchorPoint::pack (Multiplicon.cpp:210-240,
// BaseCluster.cpp:751-797, AnchorPoint.h:103-118), byte for byte.
// Two launches over the GHMs that produced multiplicons: measure (sizes + in-place
// normalisation) and write; the host scans the per-GHM sizes in between.
#include <cuda_runtime.h>

#include "common.h"
#include "workspace.h"

namespace {

constexpr int PT = 128;

struct PGhm {
    uint32_t* pool;
    ClusterRec* cl;
    MultRec* mult;
    int* out_mults;
    int* oc;      // output clusters in stream order (reuses the `act` array)
    int n_mult, n_oc;
};

__device__ void pghm_init(PGhm& p, const GhmTask& task, uint8_t* ws, int pool_factor)
{
    const int D = (int)(task.cnt[0] + task.cnt[1]);
    WsLayout L = ws_layout(D, pool_factor);
    uint8_t* base = ws + task.ws_off;
    p.pool = (uint32_t*)(base + L.pool);
    p.cl = (ClusterRec*)(base + L.cl);
    p.mult = (MultRec*)(base + L.mult);
    p.out_mults = (int*)(base + L.out_mults);
    p.oc = (int*)(base + L.act);
    p.n_mult = task.n_mult;
}

__device__ __forceinline__ void put32(uint8_t* p, int32_t v)
{
    p[0] = (uint8_t)(v);
    p[1] = (uint8_t)(v >> 8);
    p[2] = (uint8_t)(v >> 16);
    p[3] = (uint8_t)(v >> 24);
}
__device__ __forceinline__ void put64(uint8_t* p, double d)
{
    const unsigned long long v = (unsigned long long)__double_as_longlong(d);
#pragma unroll
    for (int i = 0; i < 8; i++) p[i] = (uint8_t)(v >> (8 * i));
}

// does profile element (segment n, column x) pair with dataset element yg?  (Gene.cpp:17-30)
__device__ __forceinline__ bool seg_match(const ProfileDev& pf, int n, int x, int fy, int gy)
{
    const int idx = n * pf.L + x;
    if (pf.flags[idx] & 2) return false;
    return pf.fam[idx] == fy && pf.gid[idx] != gy;
}

template <bool HL>
__global__ void __launch_bounds__(PT) pack_measure_kernel(GhmTask* tasks, const int* __restrict__ sel, int n_sel,
                                                          uint8_t* ws, DatasetDev ds, ProfileDev pf, int pool_factor)
{
    __shared__ PGhm p;
    if (blockIdx.x >= n_sel) return;
    GhmTask& task = tasks[sel[blockIdx.x]];
    const int tid = threadIdx.x;
    if (task.n_mult == 0 || task.status != 0) return;  // uniform: nothing to serialise
    if (tid == 0) {
        pghm_init(p, task, ws, pool_factor);
        int n = 0;
        for (int m = 0; m < p.n_mult; m++)
            for (int c = p.mult[p.out_mults[m]].head; c >= 0; c = p.cl[c].mnext) p.oc[n++] = c;
        p.n_oc = n;
    }
    __syncthreads();
    const int64_t yoff = ds.list_off[task.y_list];
    for (int i = tid; i < p.n_oc; i += PT) {
        ClusterRec& c = p.cl[p.oc[i]];
        uint32_t* ap = p.pool + c.off;
        int mn = KEY_Y(ap[0]), mx = mn;
        for (int k = 1; k < c.n; k++) {
            const int y = KEY_Y(ap[k]);
            mn = min(mn, y);
            mx = max(mx, y);
        }
        c.min_y = mn;
        c.max_y = mx;
        int nfalse = 0;
        if (HL) {
            // filterDuplicates: unique (x,y), ordered by x then y.  The array is x-sorted, so a
            // plain sort of the packed keys gives exactly that order.
            for (int a = 1; a < c.n; a++) {
                const uint32_t v = ap[a];
                int b = a - 1;
                while (b >= 0 && ap[b] > v) {
                    ap[b + 1] = ap[b];
                    b--;
                }
                ap[b + 1] = v;
            }
            int w = 1;
            for (int a = 1; a < c.n; a++)
                if (ap[a] != ap[w - 1]) ap[w++] = ap[a];
            c.n = w;
            for (int a = 0; a < c.n; a++) {
                const int x = KEY_X(ap[a]);
                const int yg = (int)(yoff + KEY_Y(ap[a]));
                const int fy = ds.fam[yg], gy = ds.gid[yg];
                int matches = 0;
                for (int n = 0; n < pf.S; n++) matches += seg_match(pf, n, x, fy, gy) ? 1 : 0;
                if (matches > 1) nfalse += matches - 1;
            }
        }
        c.n_false = nfalse;
        c.has_false = nfalse > 0;
    }
    __syncthreads();
    if (tid == 0) {
        int off = 0, naps = 0;
        for (int m = 0; m < p.n_mult; m++) {
            MultRec& mr = p.mult[p.out_mults[m]];
            mr.out_off = off;
            off += 33;
            for (int c = mr.head; c >= 0; c = p.cl[c].mnext) {
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
__global__ void __launch_bounds__(PT) pack_write_kernel(const GhmTask* tasks, const int* __restrict__ sel,
                                                        const int64_t* __restrict__ out_off, int n_sel, uint8_t* ws,
                                                        DatasetDev ds, ProfileDev pf, int pool_factor,
                                                        uint8_t* __restrict__ out)
{
    __shared__ PGhm p;
    if (blockIdx.x >= n_sel) return;
    const GhmTask& task = tasks[sel[blockIdx.x]];
    const int tid = threadIdx.x;
    if (tid == 0) {
        pghm_init(p, task, ws, pool_factor);
        int n = 0;
        for (int m = 0; m < p.n_mult; m++) n += p.mult[p.out_mults[m]].count;
        p.n_oc = n;
    }
    __syncthreads();
    uint8_t* dst = out + out_off[blockIdx.x];
    const int64_t yoff = ds.list_off[task.y_list];
    const int64_t xoff = HL ? 0 : ds.list_off[task.x_list];
    // multiplicon headers (Multiplicon.cpp:214-232); bounds per Multiplicon.cpp:135-193
    for (int m = tid; m < p.n_mult; m += PT) {
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
        uint8_t* q = dst + mr.out_off;
        put32(q, task.x_object_id);
        put32(q + 4, task.y_object_id);
        put32(q + 8, task.level);
        q[12] = 0;  // isRedundant
        put32(q + 13, bx);
        put32(q + 17, ex);
        put32(q + 21, by);
        put32(q + 25, ey);
        put32(q + 29, mr.count);
    }
    // clusters + anchor points
    for (int i = tid; i < p.n_oc; i += PT) {
        const ClusterRec& c = p.cl[p.oc[i]];
        const uint32_t* ap = p.pool + c.off;
        uint8_t* q = dst + c.out_off;
        put64(q, c.p);
        q[8] = c.orient;
        put64(q + 9, c.a);
        put64(q + 17, c.has_false ? 0.0 : c.b);  // addAnchorPoint(AnchorPoint&) resets b (BaseCluster.cpp:82-85)
        put64(q + 25, c.avg_x);
        put64(q + 33, c.var_x);
        put64(q + 41, c.mrss);
        put64(q + 49, c.x_end1);
        put64(q + 57, c.x_end2);
        put64(q + 65, c.y_end1);
        put64(q + 73, c.y_end2);
        q[81] = c.was_twisted;
        put32(q + 82, KEY_X(ap[0]));
        put32(q + 86, KEY_X(ap[c.n - 1]));
        put32(q + 90, c.min_y);
        put32(q + 94, c.max_y);
        put32(q + 98, c.n + c.n_false);
        q += 102;
        if (!HL) {
            for (int a = 0; a < c.n; a++) {
                const int x = KEY_X(ap[a]), y = KEY_Y(ap[a]);
                put32(q, ds.num_id[xoff + x]);
                put32(q + 4, ds.num_id[yoff + y]);
                put32(q + 8, x);
                put32(q + 12, y);
                q[16] = 1;
                q += 17;
            }
        } else {
            // multiset order after the false anchor points were inserted at upper_bound(x): per run
            // of equal x, the real APs (y ascending) then that run's false APs in discovery order
            int a = 0;
            while (a < c.n) {
                int e = a;
                const int x = KEY_X(ap[a]);
                while (e < c.n && KEY_X(ap[e]) == x) e++;
                for (int r = a; r < e; r++) {
                    const int y = KEY_Y(ap[r]);
                    const int yg = (int)(yoff + y);
                    const int fy = ds.fam[yg], gy = ds.gid[yg];
                    int gx = -1;
                    for (int n = 0; n < pf.S; n++)
                        if (seg_match(pf, n, x, fy, gy)) {
                            gx = pf.num_id[n * pf.L + x];
                            break;
                        }
                    put32(q, gx);
                    put32(q + 4, gx < 0 ? -1 : ds.num_id[yg]);
                    put32(q + 8, x);
                    put32(q + 12, y);
                    q[16] = 1;
                    q += 17;
                }
                if (c.has_false)
                    for (int r = a; r < e; r++) {
                        const int y = KEY_Y(ap[r]);
                        const int yg = (int)(yoff + y);
                        const int fy = ds.fam[yg], gy = ds.gid[yg];
                        bool seen = false;
                        for (int n = 0; n < pf.S; n++)
                            if (seg_match(pf, n, x, fy, gy)) {
                                if (seen) {
                                    put32(q, pf.num_id[n * pf.L + x]);
                                    put32(q + 4, ds.num_id[yg]);
                                    put32(q + 8, x);
                                    put32(q + 12, y);
                                    q[16] = 0;
                                    q += 17;
                                }
                                seen = true;
                            }
                    }
                a = e;
            }
        }
    }
}

}  // namespace

void launch_pack_measure(bool hl, GhmTask* tasks, const int* sel, int n_sel, uint8_t* ws, const DatasetDev& ds,
                         const ProfileDev& pf, int pool_factor, cudaStream_t st)
{
    if (n_sel == 0) return;
    if (hl) pack_measure_kernel<true><<<n_sel, PT, 0, st>>>(tasks, sel, n_sel, ws, ds, pf, pool_factor);
    else pack_measure_kernel<false><<<n_sel, PT, 0, st>>>(tasks, sel, n_sel, ws, ds, pf, pool_factor);
}

void launch_pack_write(bool hl, const GhmTask* tasks, const int* sel, const int64_t* out_off, int n_sel, uint8_t* ws,
                       const DatasetDev& ds, const ProfileDev& pf, int pool_factor, uint8_t* out, cudaStream_t st)
{
    if (n_sel == 0) return;
    if (hl) pack_write_kernel<true><<<n_sel, PT, 0, st>>>(tasks, sel, out_off, n_sel, ws, ds, pf, pool_factor, out);
    else pack_write_kernel<false><<<n_sel, PT, 0, st>>>(tasks, sel, out_off, n_sel, ws, ds, pf, pool_factor, out);
}
