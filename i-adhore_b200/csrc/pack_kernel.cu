// K3: multiplicon assembly and serialisation into the reference's wire format.
//
// Replaces GHM::setMultiplicons (GHM.cpp:671-701) and the device-computable part of
// GHMProfile::setMultiplicons (GHMProfile.cpp:68-127): Cluster::setBounds (Cluster.h:27-33),
// anchor-point gene ids, BaseCluster::filterDuplicates (BaseCluster.cpp:63-75) and the
// "false" anchor points of further matching profile segments; then
// Multiplicon::pack / BaseCluster::pack / AnchorPoint::pack (Multiplicon.cpp:210-240,
// BaseCluster.cpp:751-797, AnchorPoint.h:103-118), byte for byte.
// The measuring pass (sizes + in-place normalisation, pack_measure.cuh) runs as the tail of each GHM
// inside the clustering kernels; the host scans the per-GHM sizes; this file holds the writer.  One
// CTA per GHM, one warp per base cluster, one lane per anchor point.
#include <cuda_runtime.h>

#include "common.h"
#include "workspace.h"

#include "pack_measure.cuh"

namespace {

using pk::PGhm;
using pk::pghm_init;
using pk::seg_match;

#if defined(IADH_HOST_EMU) && defined(IADH_EMU_PT)   // tests/emu: the emulated block size of the writer
constexpr int PT = IADH_EMU_PT;
#else
constexpr int PT = 128;
#endif
constexpr int PW = PT / 32;

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
__device__ __forceinline__ void put_ap(uint8_t* q, int gx, int gy, int x, int y, int real)
{
    put32(q, gx);
    put32(q + 4, gy);
    put32(q + 8, x);
    put32(q + 12, y);
    q[16] = (uint8_t)real;
}

template <bool HL>
__global__ void __launch_bounds__(PT) pack_write_kernel(const GhmTask* tasks, const int* __restrict__ sel,
                                                        const int64_t* __restrict__ out_off, int n_sel, uint8_t* ws,
                                                        DatasetDev ds, ProfileDev pf_call, int pool_factor,
                                                        uint8_t* __restrict__ out)
{
    __shared__ PGhm p;
    __shared__ uint32_t stage[PW][140];  // per warp: 32 x 17 bytes + alignment slack
    if (blockIdx.x >= n_sel) return;
    const GhmTask& task = tasks[sel[blockIdx.x]];
    const ProfileDev& pf = profile_of(pf_call, task.prof);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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
    // clusters + anchor points: one warp per cluster, one lane per anchor point
    for (int i = warp; i < p.n_oc; i += PW) {
        const ClusterRec& c = p.cl[p.oc[i]];
        const uint32_t* ap = p.pool + c.off;
        uint8_t* q = dst + c.out_off;
        if (lane == 0) {
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
        }
        q += 102;
        if (!HL) {
            // 17-byte records at arbitrary alignment: a chunk of 32 anchor points (544 contiguous
            // output bytes) is assembled in shared memory at the same alignment phase as its
            // destination, then stored as aligned 32-bit words (byte stores only at the two edges)
            uint8_t* sb = reinterpret_cast<uint8_t*>(stage[warp]);
            for (int a0 = 0; a0 < c.n; a0 += 32) {
                const int cnt = min(32, c.n - a0);
                uint8_t* g0 = q + 17 * (size_t)a0;
                const int ph = (int)(reinterpret_cast<uintptr_t>(g0) & 3);
                if (lane < cnt) {
                    const int x = KEY_X(ap[a0 + lane]), y = KEY_Y(ap[a0 + lane]);
                    put_ap(sb + ph + 17 * lane, ds.num_id[xoff + x], ds.num_id[yoff + y], x, y, 1);
                }
                __syncwarp();
                const int nb = 17 * cnt;
                const int head = (4 - ph) & 3;                 // bytes before the first aligned word
                const int nwords = (nb - min(head, nb)) >> 2;
                const int tail0 = min(head, nb) + 4 * nwords;  // first byte after the last full word
                if (lane < min(head, nb)) g0[lane] = sb[ph + lane];
                uint32_t* gw = reinterpret_cast<uint32_t*>(g0 + head);
                const uint32_t* sw = reinterpret_cast<const uint32_t*>(sb + ph + head);
                for (int w = lane; w < nwords; w += 32) gw[w] = sw[w];
                if (lane < nb - tail0) g0[tail0 + lane] = sb[ph + tail0 + lane];
                __syncwarp();
            }
        } else {
            // multiset order after the false anchor points were inserted at upper_bound(x): per run of
            // equal x, the real APs (y ascending) then that run's false APs in discovery order.
            //   real AP a   -> a + F(run start)            F(k) = false APs of all APs before k
            //   its falses  -> run end + F(a) + 0,1,...
            const int* scr = p.scratch + c.cap;
            for (int a = lane; a < c.n; a += 32) {
                const int x = KEY_X(ap[a]), y = KEY_Y(ap[a]);
                int rs = a, re = a + 1;
                while (rs > 0 && KEY_X(ap[rs - 1]) == x) rs--;
                while (re < c.n && KEY_X(ap[re]) == x) re++;
                const int yg = (int)(yoff + y);
                const int fy = ds.fam[yg], gy = ds.gid[yg];
                const int gyn = ds.num_id[yg];
                int pos_false = re + (c.has_false ? scr[a] : 0);
                const int pos_real = a + (c.has_false ? scr[rs] : 0);
                bool seen = false;
                for (int s = 0; s < pf.S; s++)
                    if (pk::seg_match(ds, pf, s, x, yg, fy, gy)) {
                        const int gx = pf.num_id[s * pf.L + x];
                        if (!seen) put_ap(q + 17 * pos_real, gx, gyn, x, y, 1);
                        else put_ap(q + 17 * (pos_false++), gx, gyn, x, y, 0);
                        seen = true;
                    }
                if (!seen) put_ap(q + 17 * pos_real, -1, -1, x, y, 1);  // (the reference exits here)
            }
        }
    }
}

}  // namespace

#ifndef IADH_HOST_EMU
void launch_pack_write(bool hl, const GhmTask* tasks, const int* sel, const int64_t* out_off, int n_sel, uint8_t* ws,
                       const DatasetDev& ds, const ProfileDev& pf, int pool_factor, uint8_t* out, cudaStream_t st)
{
    if (n_sel == 0) return;
    if (hl) pack_write_kernel<true><<<n_sel, PT, 0, st>>>(tasks, sel, out_off, n_sel, ws, ds, pf, pool_factor, out);
    else pack_write_kernel<false><<<n_sel, PT, 0, st>>>(tasks, sel, out_off, n_sel, ws, ds, pf, pool_factor, out);
}
#endif  // IADH_HOST_EMU
