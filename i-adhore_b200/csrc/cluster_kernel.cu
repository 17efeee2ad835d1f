// K2: collinear-cluster detection inside one GHM -- one CTA per GHM, persistent grid.
//
// Replaces GHM::runCollinear (GHM.cpp:157-219) and everything below it: seedBaseClusters
// (GHM.cpp:238-324), enrichClusters (331-390), joinClusters / getClosestClusters (392-444),
// filterBaseClusters{,BF,FDR} (446-540) with BaseCluster::calculateProbability
// (BaseCluster.cpp:548-590), clusterClusters / twistClusters (542-632).
//
// The reference's passes are order-dependent greedy loops; results must be bit-identical, so
// the CTA never reorders *decisions*.  It parallelises the search inside each decision:
//   seed   : successor pointers nxt[d] (argmin-dpd dot in d's window) are computed for all dots
//            in parallel; chains are pointer walks; 256 start dots are evaluated at once and the
//            first accepted one (in dot order) is committed, then only pointers into erased
//            dots are repaired.
//   enrich : 256 dots x all active clusters evaluated at once; the first assignable dot is
//            committed, later dots of the batch re-test only the changed cluster.
//   join   : all cluster pairs evaluated in parallel, lexicographic-first minimum reduced.
// State lives in a per-GHM workspace in HBM (L2-resident for all but the densest GHMs).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#include "cluster_math.cuh"
#include "common.h"
#include "pack_measure.cuh"
#include "workspace.h"

namespace {

constexpr int MAXW = 16;  // warps per CTA, upper bound (block size is a template parameter)

// The clustering of one GHM is a dependent chain that keeps about ONE warp busy (the one holding
// "thread 0" of the serial sections).  Warp w of every CTA lands on SM sub-partition w % 4, so with
// several CTAs per SM all busy warps would share one scheduler (measured: 2x slowdown at 4 CTAs/SM).
// Each CTA therefore rotates its thread numbering by whole warps: the k-th CTA to arrive on an SM
// puts its virtual thread 0 in warp k % NW.  All code below uses the virtual id VTID.
__shared__ int s_rot;
static __device__ __forceinline__ unsigned read_smid()
{
#ifdef IADH_HOST_EMU   // tests/emu: the kernels compiled for the host, lanes as fibers
    return 0;
#else
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    return smid;
#endif
}
#define VTID ((int)((threadIdx.x + (unsigned)s_rot) & (unsigned)(NT - 1)))

struct Ghm {
    uint32_t* key[2];
    uint8_t* alive[2];
    int nk[2];
    int* nxt;
    uint32_t* pool;
    uint8_t* pflag;
    // The two orientations of a GHM are clustered independently up to the final enrichment
    // (GHM.cpp:167-184 touches matrix[o] / baseclusters[o] only), so they run on two CTAs.  Each owns
    // a region of the AP pool and of the cluster table; region r = orientation r.
    int pool_lo[2], pool_used[2], pool_hi[2];   // absolute indices into pool
    ClusterRec* cl;
    int cl_lo[2], ncl[2], cl_hi[2];             // cluster ids of region r: [cl_lo[r], ncl[r])
    int* bc[2];
    int nbc[2];
    int* act;
    int* act_next;
    uint8_t* changed;
    MultRec* mult;
    int nmult;
    int* out_mults;
    double* rb_d;   // join cache: best distance / partner id of every cluster, rank, redo list
    int* rb_j;
    int* rank;
    int* pos;       // join: current position of a cluster id in its orientation vector
    int* redo;
    int2* xr;       // enrich: per active cluster [x_end1 - gap, x_end2 + gap] (exact prefilter, dpd >= |dx|)
    double* dscr;   // 2 doubles per anchor point: terms of the ordered mrss / var_x sums
    uint32_t* tmp;
    uint8_t* tmpf;  // staging flags, parallel to tmp
    int scratch_cap;  // entries of tmp / doubles of dscr usable as whole-GHM scratch (phase 2)
    // capacities of the view (the global workspace is sized by worst-case bounds, the shared-memory
    // resident view by budgets: exceeding one sets `status`, the host re-runs the GHM on the global path)
    int key_cap[2];   // entries of key[o] / alive[o]
    int tmp_cap;      // entries of tmp / tmpf
    int dscr_cap;     // anchor points per chunk of the ordered statistics sums (2 doubles each)
    int tight;        // resident view: tight anchor-point block capacities, cluster slots recycled after merges
    int free_head[2]; // recycled cluster slots per orientation (linked through ClusterRec::mnext), -1 = none
    // seeding state per orientation (seed_pass): the successor pointers are gap-independent and persist
    // across the gap sizes; g3 = smallest gap size at which some 3-dot chain exists (exact while no dot
    // of the matrix was erased since it was computed: `dirty`), so seed passes below it are skipped
    int nxt_off[2];   // start of orientation o's successor pointers inside nxt
    int seeded[2], dirty[2], g3[2];
    long long* ctr;   // measurement aid (resident kernel under IADH_PASS_TIMING): event counters, may be null
    int status;
    int coop[4];
    // task constants
    uint32_t cnt[2];
    double area;
    int level;
};

struct Red {
    double d[MAXW];
    long long k[MAXW];
    int i[MAXW];
};

// Everything below is parameterised on the CTA size NT: small GHMs run on 32- or 128-thread CTAs
// (more GHMs in flight per SM, barriers almost free), heavy ones on 256 threads.
// CP ("compact"): the big inlined helpers become single out-of-line copies.  The pass-per-launch
// kernels are instruction-fetch bound (32 single-warp CTAs per SM), so a smaller hot image is worth
// the call overhead there; the fused kernel keeps everything inline (latency first).
template <int NT, bool CP = false>
struct KT {
static constexpr int NW = NT / 32;

// ---- block reductions -------------------------------------------------------------------

// minimum int over the block (INT_MAX = none); two barriers
static __device__ int block_min_int(int v, int* sh)
{
    const int lane = VTID & 31, w = VTID >> 5;
    v = __reduce_min_sync(0xffffffffu, v);  // REDUX: one instruction, signed min
    if (lane == 0) sh[w] = v;
    __syncthreads();
    int r = sh[0];
#pragma unroll
    for (int i = 1; i < NW; i++) r = min(r, sh[i]);
    __syncthreads();
    return r;
}

// lexicographic minimum of (dist, idx); idx < 0 = none
static __device__ void block_min_pair(double& d, long long& k, Red* sh)
{
    const int lane = VTID & 31, w = VTID >> 5;
#pragma unroll
    for (int s = 16; s; s >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, d, s);
        const long long ok = __shfl_xor_sync(0xffffffffu, k, s);
        if (ok >= 0 && (k < 0 || od < d || (od == d && ok < k))) {
            d = od;
            k = ok;
        }
    }
    if (lane == 0) {
        sh->d[w] = d;
        sh->k[w] = k;
    }
    __syncthreads();
    d = sh->d[0];
    k = sh->k[0];
#pragma unroll
    for (int i = 1; i < NW; i++) {
        const double od = sh->d[i];
        const long long ok = sh->k[i];
        if (ok >= 0 && (k < 0 || od < d || (od == d && ok < k))) {
            d = od;
            k = ok;
        }
    }
    __syncthreads();
}

// ---- AP pool ------------------------------------------------------------------------------

// Anchor-point pool: bump allocation of self-describing blocks [owner cluster, capacity,
// payload...].  Growth of a cluster abandons its old block; when the pool runs out, the live
// blocks (those a live cluster still points at) are slid down in allocation order and trimmed
// to the capacity class of their current size.  Live data is at most ~4.7 D entries, so with
// pool_factor >= 8 an allocation can only fail after a collection if the host under-sized the
// pool on purpose (then the GHM reports status 1 and the engine reruns it with a larger pool).
static __device__ __forceinline__ int next_cap(int n)
{
    int c = 8;
    while (c < n) c <<= 1;
    return c;
}
// block capacity policies: the global pool (8 D entries) rounds to powers of two; the resident pool (2 D
// entries in shared memory) allocates what is needed and grows by half
static __device__ __forceinline__ int cap_for(const Ghm& g, int n) { return g.tight ? ((n + 3) & ~3) : next_cap(n); }
static __device__ __forceinline__ int cap_grow(const Ghm& g, int cap) { return g.tight ? cap + max(4, (cap >> 1) & ~3) : cap * 2; }
static __device__ __forceinline__ int cap_seed(const Ghm& g, int len) { return g.tight ? ((len + 2 + 3) & ~3) : next_cap(len + (len >> 1)); }

static __device__ __forceinline__ bool valid_id(const Ghm& g, int id)
{
    return (id >= g.cl_lo[0] && id < g.ncl[0]) || (id >= g.cl_lo[1] && id < g.ncl[1]);
}

static __device__ void pool_gc_body(Ghm& g, int rg)
{
    int r = g.pool_lo[rg], w = r;
    while (r < g.pool_used[rg]) {
        const int owner = (int)g.pool[r];
        const int cap = (int)g.pool[r + 1];
        const int payload = r + 2;
        ClusterRec& c = g.cl[owner];
        if (valid_id(g, owner) && !c.dead && c.off == payload) {
            int ncap = cap_for(g, c.n);
            if (ncap > cap) ncap = cap;
            if (w != r)
                for (int i = 0; i < c.n; i++) {
                    g.pool[w + 2 + i] = g.pool[payload + i];
                    g.pflag[w + 2 + i] = g.pflag[payload + i];
                }
            g.pool[w] = (uint32_t)owner;
            g.pool[w + 1] = (uint32_t)ncap;
            c.off = w + 2;
            c.cap = ncap;
            w += 2 + ncap;
        }
        r += 2 + cap;
    }
    g.pool_used[rg] = w;
}

#define CTR_ADD(k, v) do { if (g.ctr) g.ctr[k] += (v); } while (0)
static __device__ void pool_gc(Ghm& g, int rg)
{
    const long long t0_ = g.ctr ? clock64() : 0;
    pool_gc_body(g, rg);
    if (g.ctr) {
        g.ctr[8]++;
        g.ctr[9] += clock64() - t0_;
    }
}

// single-thread contexts only (the rest of the CTA waits at a barrier).  `rg` = preferred region
// (the cluster's orientation); the other region is a fallback, empty while the other orientation
// is still being clustered by its own CTA.
static __device__ int pool_alloc(Ghm& g, int n, int owner, int rg)
{
    for (int t = 0; t < 2; t++, rg ^= 1) {
        if (g.pool_used[rg] + n + 2 > g.pool_hi[rg]) {
            if (g.pool_hi[rg] == g.pool_lo[rg]) continue;
            pool_gc(g, rg);
            if (g.pool_used[rg] + n + 2 > g.pool_hi[rg]) continue;
        }
        const int off = g.pool_used[rg];
        g.pool[off] = (uint32_t)owner;
        g.pool[off + 1] = (uint32_t)n;
        g.pool_used[rg] += n + 2;
        return off + 2;
    }
    g.status = 1;
    return -1;
}


// multiset<AnchorPoint>::insert (BaseCluster.cpp:58-61): after the last AP with x <= X
static __device__ bool insert_ap(Ghm& g, ClusterRec& c, uint32_t key)
{
    if (c.n == c.cap) {
        const int ncap = cap_grow(g, c.cap);
        const int noff = pool_alloc(g, ncap, (int)(&c - g.cl), c.orient);
        if (noff < 0) return false;
        for (int i = 0; i < c.n; i++) {
            g.pool[noff + i] = g.pool[c.off + i];
            g.pflag[noff + i] = g.pflag[c.off + i];
        }
        c.off = noff;
        c.cap = ncap;
    }
    const int x = KEY_X(key);
    int lo = 0, hi = c.n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (KEY_X(g.pool[c.off + mid]) <= x) lo = mid + 1;
        else hi = mid;
    }
    for (int i = c.n; i > lo; i--) {
        g.pool[c.off + i] = g.pool[c.off + i - 1];
        g.pflag[c.off + i] = g.pflag[c.off + i - 1];
    }
    g.pool[c.off + lo] = key;
    g.pflag[c.off + lo] = 0;
    c.n++;
    const long long X = x, Y = KEY_Y(key);
    c.sx += X; c.sy += Y; c.sxy += X * Y; c.sxx += X * X; c.syy += Y * Y;
    return true;
}

// BaseCluster::mergeWith (BaseCluster.cpp:338-351): B's APs are inserted one by one, i.e. a
// stable merge with A's APs first on equal x; B's backbone points are dropped where A already
// has a backbone point with the same x (set<AnchorPoint> on x).
static __device__ bool merge_clusters(Ghm& g, ClusterRec& A, const ClusterRec& B)
{
    const int n = A.n + B.n;
    if (n > A.cap) {
        const int ncap = cap_for(g, n);
        const int noff = pool_alloc(g, ncap, (int)(&A - g.cl), A.orient);
        if (noff < 0) return false;
        for (int i = 0; i < A.n; i++) {
            g.pool[noff + i] = g.pool[A.off + i];
            g.pflag[noff + i] = g.pflag[A.off + i];
        }
        A.off = noff;
        A.cap = ncap;
    }
    int ia = A.n - 1, ib = B.n - 1, k = n - 1;
    while (ib >= 0) {
        const uint32_t kb = g.pool[B.off + ib];
        if (ia >= 0 && KEY_X(g.pool[A.off + ia]) > KEY_X(kb)) {
            g.pool[A.off + k] = g.pool[A.off + ia];
            g.pflag[A.off + k] = g.pflag[A.off + ia];
            ia--;
        } else {
            uint8_t fl = g.pflag[B.off + ib];
            if (fl) {
                for (int j = ia; j >= 0 && KEY_X(g.pool[A.off + j]) == KEY_X(kb); j--)
                    if (g.pflag[A.off + j]) {
                        fl = 0;
                        break;
                    }
            }
            g.pool[A.off + k] = kb;
            g.pflag[A.off + k] = fl;
            ib--;
        }
        k--;
    }
    A.n = n;
    A.sx += B.sx; A.sy += B.sy; A.sxy += B.sxy; A.sxx += B.sxx; A.syy += B.syy;
    cm::update_stats(A, g.pool);
    return true;
}

// ---- cooperative (whole-CTA) cluster updates ----------------------------------------------------
// The reference recomputes a cluster's statistics from scratch after every change, with mrss and
// var_x accumulated sequentially in multiset order.  That order is kept: all threads compute the
// per-AP terms, one thread adds them up in order (a pure dependent-add chain, no loads or
// multiplies on it).  Short clusters stay on one thread.
#define COOP_T(k) (g.coop[k])

static __device__ void stats_coop_il(Ghm& g, ClusterRec& c)
{
    const int tid = VTID;
    const int n = c.n;
    if (n < COOP_T(0)) {
        if (tid == 0) cm::update_stats(c, g.pool);
        __syncthreads();
        return;
    }
    if (tid == 0) {
        cm::regress(c);
        CTR_ADD(14, n);
    }
    __syncthreads();
    // the per-AP terms are staged chunk by chunk (the global view holds the whole cluster in one chunk, the
    // resident view dscr_cap anchor points); thread 0 carries the two running sums across chunks
    double mrss = 0.0, var_x = 0.0;
    const int chunk = g.dscr_cap;
    for (int base = 0; base < n; base += chunk) {
        const int m = min(chunk, n - base);
        {
            const double a = c.a, b = c.b, avg_x = c.avg_x;
            const uint32_t* ap = g.pool + c.off + base;
            for (int i = tid; i < m; i += NT) {
                const uint32_t k = ap[i];
                const int x = KEY_X(k), y = KEY_Y(k);
                const double y_est = b * x + a;
                g.dscr[2 * i] = (y - y_est) * (y - y_est);
                g.dscr[2 * i + 1] = (x - avg_x) * (x - avg_x);
            }
        }
        __syncthreads();
        if (tid == 0) {
            const double2* t = reinterpret_cast<const double2*>(g.dscr);
            int i = 0;
            for (; i + 8 <= m; i += 8) {
                double2 v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) v[u] = t[i + u];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    mrss += v[u].x;
                    var_x += v[u].y;
                }
            }
            for (; i < m; i++) {
                mrss += t[i].x;
                var_x += t[i].y;
            }
        }
        if (base + chunk < n) __syncthreads();   // the staging area is reused by the next chunk
    }
    if (tid == 0) cm::finish_stats(c, mrss, var_x, g.pool);
    __syncthreads();
}
static __device__ __noinline__ void stats_coop_ni(Ghm& g, ClusterRec& c) { stats_coop_il(g, c); }
static __device__ __forceinline__ void stats_coop(Ghm& g, ClusterRec& c)
{
    if (CP) stats_coop_ni(g, c);
    else stats_coop_il(g, c);
}

static __device__ __forceinline__ int lower_bound_x(const uint32_t* a, int n, int x)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (KEY_X(a[mid]) < x) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}
static __device__ __forceinline__ int upper_bound_x(const uint32_t* a, int n, int x)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (KEY_X(a[mid]) <= x) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// BaseCluster::mergeWith by the whole CTA: every element finds its final position by a binary
// search into the other cluster (A's APs first on equal x), staged, then copied.
static __device__ bool merge_coop(Ghm& g, int idA, int idB)
{
    const int tid = VTID;
    ClusterRec& A = g.cl[idA];
    const ClusterRec& B = g.cl[idB];
    const int nA = A.n, nB = B.n, n = nA + nB;
    if (n < COOP_T(1)) {
        __syncthreads();
        if (tid == 0) merge_clusters(g, A, B);
        __syncthreads();
        return g.status == 0;
    }
    if (n > g.tmp_cap) {   // (resident view only: the staging area is a budget)
        __syncthreads();
        if (tid == 0) g.status = 3;
        __syncthreads();
        return false;
    }
    uint32_t* skey = g.tmp;
    uint8_t* sflag = g.tmpf;
    {
        const uint32_t* pa = g.pool + A.off;
        const uint32_t* pb = g.pool + B.off;
        const uint8_t* fa = g.pflag + A.off;
        const uint8_t* fb = g.pflag + B.off;
        for (int i = tid; i < nA; i += NT) {
            const int pos = i + lower_bound_x(pb, nB, KEY_X(pa[i]));
            skey[pos] = pa[i];
            sflag[pos] = fa[i];
        }
        for (int j = tid; j < nB; j += NT) {
            const int x = KEY_X(pb[j]);
            const int ub = upper_bound_x(pa, nA, x);
            uint8_t fl = fb[j];
            if (fl)
                for (int k = ub - 1; k >= 0 && KEY_X(pa[k]) == x; k--)
                    if (fa[k]) {
                        fl = 0;
                        break;
                    }
            skey[j + ub] = pb[j];
            sflag[j + ub] = fl;
        }
    }
    __syncthreads();
    if (tid == 0) {
        bool ok = true;
        if (n > A.cap) {
            const int ncap = cap_for(g, n);
            const int noff = pool_alloc(g, ncap, idA, A.orient);  // may compact the pool: A/B blocks can move
            if (noff >= 0) {
                A.off = noff;
                A.cap = ncap;
            } else
                ok = false;
        }
        if (ok) {
            A.n = n;
            A.sx += B.sx; A.sy += B.sy; A.sxy += B.sxy; A.sxx += B.sxx; A.syy += B.syy;
        }
    }
    __syncthreads();
    if (g.status) return false;
    for (int i = tid; i < n; i += NT) {
        g.pool[A.off + i] = skey[i];
        g.pflag[A.off + i] = sflag[i];
    }
    __syncthreads();
    stats_coop(g, A);
    return true;
}

// BaseCluster::addAnchorPoint(x, y) + updateStatistics by the whole CTA
static __device__ bool insert_coop(Ghm& g, ClusterRec& c, uint32_t key)
{
    const int tid = VTID;
    const int n = c.n;
    if (n < COOP_T(2)) {
        if (tid == 0 && insert_ap(g, c, key)) cm::update_stats(c, g.pool);
        __syncthreads();
        return g.status == 0;
    }
    if (tid == 0) CTR_ADD(12, 1);
    if (tid == 0 && n == c.cap) {
        CTR_ADD(10, n);
        const int ncap = cap_grow(g, c.cap);
        const int noff = pool_alloc(g, ncap, (int)(&c - g.cl), c.orient);
        if (noff >= 0) {
            for (int i = 0; i < n; i++) {
                g.pool[noff + i] = g.pool[c.off + i];
                g.pflag[noff + i] = g.pflag[c.off + i];
            }
            c.off = noff;
            c.cap = ncap;
        }
    }
    __syncthreads();
    if (g.status) return false;
    uint32_t* ap = g.pool + c.off;
    uint8_t* fl = g.pflag + c.off;
    const int pos = upper_bound_x(ap, n, KEY_X(key));
    for (int base = n - 1; base >= pos; base -= NT) {
        const int i = base - tid;
        uint32_t v = 0;
        uint8_t f = 0;
        if (i >= pos) {
            v = ap[i];
            f = fl[i];
        }
        __syncthreads();
        if (i >= pos) {
            ap[i + 1] = v;
            fl[i + 1] = f;
        }
        __syncthreads();
    }
    if (tid == 0) {
        ap[pos] = key;
        fl[pos] = 0;
        c.n = n + 1;
        const long long X = KEY_X(key), Y = KEY_Y(key);
        c.sx += X; c.sy += Y; c.sxy += X * Y; c.sxx += X * X; c.syy += Y * Y;
    }
    __syncthreads();
    stats_coop(g, c);
    return true;
}

// ---- seed ---------------------------------------------------------------------------------

static __device__ __forceinline__ int lower_bound_key(const uint32_t* a, int lo, int hi, uint32_t v)
{
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// GHM::seedBaseCluster inner search (GHM.cpp:248-276): closest alive dot in the window, strict
// '<' so the first in (x,y) order wins ties; only dpd <= gap qualifies.
static __device__ int window_next(const uint32_t* key, const uint8_t* alive, int n, int i,
                           int gap, int orient)
{
    const uint32_t k = key[i];
    const int rx = KEY_X(k), ry = KEY_Y(k);
    const int loY = orient ? ry + 1 : ry - gap;
    const int hiY = orient ? ry + gap : ry - 1;
    const int hiX = rx + gap;
    int best = gap + 1, bj = -1;
    // first dot of column rx+1 or later: the rest of column rx is usually 0-2 dots, so step over it
    // before falling back to a binary search (dense GHMs)
    int j = i + 1;
    {
        int steps = 0;
        while (j < n && KEY_X(key[j]) == rx && steps < 3) {
            j++;
            steps++;
        }
        if (j < n && KEY_X(key[j]) == rx) j = lower_bound_key(key, j + 1, n, MAKE_KEY(rx + 1, 0));
    }
    for (; j < n; j++) {
        const uint32_t kj = key[j];
        const int x = KEY_X(kj);
        if (x > hiX) break;
        const int y = KEY_Y(kj);
        if (y < loY || y > hiY || !alive[j]) continue;
        const int d = cm::dpd_i(rx, ry, x, y);
        if (d < best) {
            best = d;
            bj = j;
        }
    }
    return bj;
}

// GHM::seedBaseClusters (GHM.cpp:287-324).
// The successor of a dot at gap size g -- the first dot of minimal dpd in its g x g window, if that
// dpd <= g (GHM.cpp:248-276) -- does not depend on g: dpd >= max(dx, dy), so every dot with dpd <= g lies
// inside the g-window, and the minimal-dpd dots of the largest window (ties included, same order) are the
// minimal-dpd dots of every smaller window that holds one.  nxt[] therefore keeps the successor for the
// largest gap size `gmax` for the whole gap-size loop; a pass at gap g follows a pointer only if its dpd
// is <= g.  Pointers into dots erased since (by an accepted seed: repaired at once; by the enrichment:
// checked at the start of the next pass) are recomputed.  g3 = the smallest g for which some alive dot has
// two links <= g in a row; a pass with gap < g3 cannot build a 3-dot chain and is skipped by the callers
// (exact as long as no dot was erased after g3 was computed: erasures can shorten a chain's links).
static __device__ __forceinline__ int succ_at(const uint32_t* key, const int* nxt, int i, int gap)
{
    const int j = nxt[i];
    if (j < 0) return -1;
    const uint32_t a = key[i], b = key[j];
    return cm::dpd_i(KEY_X(a), KEY_Y(a), KEY_X(b), KEY_Y(b)) <= gap ? j : -1;
}

static __device__ void seed_pass(Ghm& g, int gap, int gmax, int o, double q, int* sh_i, int* sh_s)
{
    const int tid = VTID;
    const uint32_t* key = g.key[o];
    uint8_t* alive = g.alive[o];
    const int n = g.nk[o];
    int* nxt = g.nxt + g.nxt_off[o];
    if (!g.seeded[o]) {
        for (int i = tid; i < n; i += NT) nxt[i] = alive[i] ? window_next(key, alive, n, i, gmax, o) : -1;
    } else {
        for (int i = tid; i < n; i += NT)
            if (alive[i]) {
                const int j = nxt[i];
                if (j >= 0 && !alive[j]) nxt[i] = window_next(key, alive, n, i, gmax, o);
            }
    }
    __syncthreads();
    int base = 0;
    while (base < n) {
        const int i = base + tid;
        bool acc = false;
        if (i < n && alive[i]) {
            const int j1 = succ_at(key, nxt, i, gap);
            if (j1 >= 0 && succ_at(key, nxt, j1, gap) >= 0) {
                long long sx = 0, sy = 0, sxy = 0, sxx = 0, syy = 0;
                unsigned len = 0;
                for (int j = i; j >= 0; j = succ_at(key, nxt, j, gap)) {
                    const uint32_t k = key[j];
                    const long long x = KEY_X(k), y = KEY_Y(k);
                    sx += x; sy += y; sxy += x * y; sxx += x * x; syy += y * y;
                    len++;
                }
                acc = cm::r2_sums(sx, sy, sxy, sxx, syy, len) >= q;  // len > 2 by construction
            }
        }
        const int first = block_min_int(acc ? tid : 0x7fffffff, sh_i);
        if (first == 0x7fffffff) {
            base += NT;
            continue;
        }
        if (tid == first) {
            // commit: new base cluster from the chain (GHM.cpp:301-311)
            int len = 0;
            for (int j = i; j >= 0; j = succ_at(key, nxt, j, gap)) len++;
            const int cap = cap_seed(g, len);
            // a slot freed by a merge (resident view) or the next fresh one
            const bool reuse = g.free_head[o] >= 0;
            const int nid = reuse ? g.free_head[o] : g.ncl[o];
            const int off = (nid < g.cl_hi[o]) ? pool_alloc(g, cap, nid, o) : -1;
            if (off >= 0) {
                if (reuse) g.free_head[o] = g.cl[nid].mnext;
                CTR_ADD(13, 1);
                ClusterRec& c = g.cl[nid];
                c.off = off; c.cap = cap; c.n = len;
                c.mult = -1; c.mnext = -1; c.orient = (unsigned char)o; c.was_twisted = 0; c.dead = 0;
                c.a = 0; c.b = 0; c.avg_x = 0; c.p = 0;
                int w = 0, lastj = i;
                for (int j = i; j >= 0;) {
                    const int jn = succ_at(key, nxt, j, gap);   // (before the dot is erased: same pointer either way)
                    g.pool[off + w] = key[j];
                    g.pflag[off + w] = 1;
                    alive[j] = 0;
                    lastj = j;
                    w++;
                    j = jn;
                }
                cm::recompute_sums(c, g.pool);
                g.bc[o][g.nbc[o]++] = nid;
                if (!reuse) g.ncl[o] = nid + 1;
                sh_s[0] = lastj;
                sh_s[6] = nid;
            } else {
                if (!g.status) g.status = 2;  // cluster table full
                sh_s[0] = -1;
            }
        }
        __syncthreads();
        if (g.status) return;
        stats_coop(g, g.cl[sh_s[6]]);
        {
            // repair successor pointers that point into the erased chain (any dot within gmax columns before it)
            const int lastj = sh_s[0];
            const int cx = KEY_X(key[base + first]);
            const int lo = lower_bound_key(key, 0, n, MAKE_KEY(max(cx - gmax, 0), 0));
            for (int j = lo + tid; j < lastj; j += NT)
                if (alive[j] && nxt[j] >= 0 && !alive[nxt[j]]) nxt[j] = window_next(key, alive, n, j, gmax, o);
        }
        base = base + first + 1;
        __syncthreads();
    }
    // smallest gap size with a 3-dot chain among the dots left (every pointer of an alive dot is current here)
    int m = 0x7fffffff;
    for (int i = tid; i < n; i += NT)
        if (alive[i]) {
            const int j = nxt[i];
            if (j >= 0) {
                const int k = nxt[j];
                if (k >= 0) {
                    const uint32_t a = key[i], b = key[j], c = key[k];
                    m = min(m, max(cm::dpd_i(KEY_X(a), KEY_Y(a), KEY_X(b), KEY_Y(b)), cm::dpd_i(KEY_X(b), KEY_Y(b), KEY_X(c), KEY_Y(c))));
                }
            }
        }
    m = block_min_int(m, sh_i);
    if (tid == 0) {
        g.g3[o] = m;
        g.seeded[o] = 1;
        g.dirty[o] = 0;
    }
    __syncthreads();
}
// a seed pass at this gap size can be skipped: no 3-dot chain exists below g3 (see seed_pass)
static __device__ __forceinline__ bool seed_skippable(const Ghm& g, int gap, int o)
{
    return g.seeded[o] && !g.dirty[o] && gap < g.g3[o];
}

// ---- enrich -------------------------------------------------------------------------------

// one (dot, cluster) test of GHM::enrichClusters (GHM.cpp:352-365)
static __device__ __forceinline__ bool enrich_test_il(const ClusterRec& c, int x, int y, double limit, double q, double& dist)
{
    dist = cm::dist_to_point(c, x, y);
    if (dist >= limit) return false;
    return cm::r2_with_point(c, x, y) >= q && cm::in_interval(c, x, y);
}
static __device__ __noinline__ bool enrich_test_ni(const ClusterRec& c, int x, int y, double limit, double q, double& dist)
{
    return enrich_test_il(c, x, y, limit, q, dist);
}
static __device__ __forceinline__ bool enrich_test(const ClusterRec& c, int x, int y, double limit, double q, double& dist)
{
    return CP ? enrich_test_ni(c, x, y, limit, q, dist) : enrich_test_il(c, x, y, limit, q, dist);
}

// all active clusters against one dot, in vector order (GHM.cpp:352-365).  g.xr[k] holds the
// cluster's x-range widened by the gap: outside it distanceToPoint >= |dx| >= gap+1 >= bd, i.e. the
// reference's own `distance < closest` test fails -- the prefilter rejects exactly those pairs.
static __device__ __forceinline__ void enrich_scan_il(const Ghm& g, const int* act, int nact, int x, int y, double q,
                                                   double& bd, int& bk)
{
    const int2* xr = g.xr;
    int k = 0;
    for (; k + 4 <= nact; k += 4) {
        const int2 r0 = xr[k], r1 = xr[k + 1], r2 = xr[k + 2], r3 = xr[k + 3];
        unsigned m = (x >= r0.x && x <= r0.y ? 1u : 0u) | (x >= r1.x && x <= r1.y ? 2u : 0u) |
                     (x >= r2.x && x <= r2.y ? 4u : 0u) | (x >= r3.x && x <= r3.y ? 8u : 0u);
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            double d;
            if (enrich_test(g.cl[act[k + b]], x, y, bd, q, d)) {
                bd = d;
                bk = k + b;
            }
        }
    }
    for (; k < nact; k++) {
        const int2 r = xr[k];
        if (x < r.x || x > r.y) continue;
        double d;
        if (enrich_test(g.cl[act[k]], x, y, bd, q, d)) {
            bd = d;
            bk = k;
        }
    }
}
static __device__ __noinline__ void enrich_scan_ni(const Ghm& g, const int* act, int nact, int x, int y, double q, double& bd, int& bk)
{
    enrich_scan_il(g, act, nact, x, y, q, bd, bk);
}
static __device__ __forceinline__ void enrich_scan(const Ghm& g, const int* act, int nact, int x, int y, double q, double& bd, int& bk)
{
    if (CP) enrich_scan_ni(g, act, nact, x, y, q, bd, bk);
    else enrich_scan_il(g, act, nact, x, y, q, bd, bk);
}

static __device__ void enrich_pass(Ghm& g, int gap, int oc, int om, double q, int* sh_i, int* sh_s)
{
    const int tid = VTID;
    const uint32_t* key = g.key[om];
    uint8_t* alive = g.alive[om];
    const int n = g.nk[om];
    int nact = g.nbc[oc];
    for (int i = tid; i < nact; i += NT) g.act[i] = g.bc[oc][i];
    __syncthreads();
    int* act = g.act;
    int* act_next = g.act_next;
    while (nact > 0) {
        for (int i = tid; i < nact; i += NT) {
            g.changed[i] = 0;
            const ClusterRec& c = g.cl[act[i]];
            g.xr[i] = make_int2((int)c.x_end1 - gap, (int)c.x_end2 + gap);
        }
        if (tid == 0) sh_s[1] = 0;  // number of clusters in the next round
        __syncthreads();
        for (int base = 0; base < n; base += NT) {
            const int i = base + tid;
            bool mine = (i < n) && alive[i];
            int x = 0, y = 0;
            double bd = gap + 1;
            int bk = -1;
            if (mine) {
                x = KEY_X(key[i]);
                y = KEY_Y(key[i]);
                enrich_scan(g, act, nact, x, y, q, bd, bk);
            }
            int from = 0;
            while (true) {
                const int first = block_min_int((bk >= 0 && tid >= from) ? tid : 0x7fffffff, sh_i);
                if (first == 0x7fffffff) break;
                if (tid == first) {
                    sh_s[0] = bk;
                    sh_s[3] = i;
                    bk = -1;
                    mine = false;
                }
                __syncthreads();
                const int kc = sh_s[0];
                {
                    const int di = sh_s[3];
                    const bool ok = insert_coop(g, g.cl[act[kc]], key[di]);
                    if (tid == 0 && ok) {
                        if (!g.changed[kc]) {
                            g.changed[kc] = 1;
                            act_next[sh_s[1]++] = act[kc];
                        }
                        alive[di] = 0;
                        g.dirty[om] = 1;   // (seed_pass: the 3-dot-chain bound of this matrix is no longer exact)
                        const ClusterRec& c = g.cl[act[kc]];
                        g.xr[kc] = make_int2((int)c.x_end1 - gap, (int)c.x_end2 + gap);
                    }
                    __syncthreads();
                }
                if (g.status) return;
                if (mine && tid > first) {
                    if (bk == kc) {
                        bd = gap + 1;
                        bk = -1;
                        enrich_scan(g, act, nact, x, y, q, bd, bk);
                    } else {
                        double d;
                        // the changed cluster wins if strictly closer, or equally close and earlier
                        const double limit = (bk >= 0 && kc < bk) ? nextafter(bd, 1e300) : bd;
                        if (enrich_test(g.cl[act[kc]], x, y, limit, q, d)) {
                            bd = d;
                            bk = kc;
                        }
                    }
                }
                from = first + 1;
                // (block_min_int's barriers order these reads against the next commit)
            }
        }
        __syncthreads();
        nact = sh_s[1];
        int* t = act;
        act = act_next;
        act_next = t;
        __syncthreads();
    }
}

// ---- join / closest clusters ---------------------------------------------------------------

// GHM::getClosestClusters (GHM.cpp:406-444): first (i,j) in lexicographic order attaining the
// minimum distance < gap+1 among pairs passing the interval and r^2 tests.
static __device__ bool closest_clusters(Ghm& g, int gap, int oi, int oj, double q, int& pi, int& pj, Red* red)
{
    const int ni = g.nbc[oi], nj = g.nbc[oj];
    const bool same = (oi == oj);
    const long long total = (long long)ni * nj;
    double bd = gap + 1;
    long long bk = -1;
    for (long long idx = VTID; idx < total; idx += NT) {
        const int i = (int)(idx / nj), j = (int)(idx - (long long)i * nj);
        if (same && j <= i) continue;
        const ClusterRec& cI = g.cl[g.bc[oi][i]];
        const ClusterRec& cJ = g.cl[g.bc[oj][j]];
        if (cI.mult >= 0 && cI.mult == cJ.mult) continue;
        // exact cheap rejection: every candidate distance of distanceToCluster is a dpd between a
        // point with x in A's x-range and one with x in B's, and dpd >= |dx|
        const double xgap = cm::maxd(cI.x_end1 - cJ.x_end2, cJ.x_end1 - cI.x_end2);
        if (xgap >= bd) continue;
        const double d = cm::dist_to_cluster(cI, cJ, g.pool[cI.off], g.pool[cI.off + cI.n - 1], g.pool[cJ.off],
                                             g.pool[cJ.off + cJ.n - 1]);
        if (d < bd) {
            if (cm::r2_merged(cI, cJ) >= q && cm::partial_overlap(cI, cJ, g.pool)) {
                bd = d;
                bk = idx;
            }
        }
    }
    block_min_pair(bd, bk, red);
    if (bk < 0) return false;
    pi = (int)(bk / nj);
    pj = (int)(bk - (long long)pi * nj);
    return true;
}

// One candidate pair of getClosestClusters (GHM.cpp:420-438): distance, then the interval and
// merged-r^2 tests.  `bd`/`brank` = best so far of the caller; a pair replaces it when closer,
// or equally close and earlier in vector order (the reference scans pairs lexicographically
// with a strict '<').  Rejections by x / y extent are exact: every candidate distance of
// distanceToCluster is a dpd between a point of A's hull and one of B's, and dpd >= |dx|, |dy|.
static __device__ __forceinline__ bool pair_better_il(const Ghm& g, const ClusterRec& cI, const ClusterRec& cJ, double q,
                                            double bd, int brank, int rankJ, double& dout)
{
    const double xgap = cm::maxd(cI.x_end1 - cJ.x_end2, cJ.x_end1 - cI.x_end2);
    if (xgap > bd) return false;
    const uint32_t s1 = g.pool[cI.off], e1 = g.pool[cI.off + cI.n - 1];
    const uint32_t s2 = g.pool[cJ.off], e2 = g.pool[cJ.off + cJ.n - 1];
    const double ilo = cm::mind(cm::mind(cI.y_end1, cI.y_end2), (double)min(KEY_Y(s1), KEY_Y(e1)));
    const double ihi = cm::maxd(cm::maxd(cI.y_end1, cI.y_end2), (double)max(KEY_Y(s1), KEY_Y(e1)));
    const double jlo = cm::mind(cm::mind(cJ.y_end1, cJ.y_end2), (double)min(KEY_Y(s2), KEY_Y(e2)));
    const double jhi = cm::maxd(cm::maxd(cJ.y_end1, cJ.y_end2), (double)max(KEY_Y(s2), KEY_Y(e2)));
    if (cm::maxd(ilo - jhi, jlo - ihi) > bd) return false;
    const double d = cm::dist_to_cluster(cI, cJ, s1, e1, s2, e2);
    if (!(d < bd || (d == bd && rankJ < brank))) return false;
    if (!(cm::r2_merged(cI, cJ) >= q && cm::partial_overlap(cI, cJ, g.pool))) return false;
    dout = d;
    return true;
}
static __device__ __noinline__ bool pair_better_ni(const Ghm& g, const ClusterRec& cI, const ClusterRec& cJ, double q, double bd,
                                                   int brank, int rankJ, double& dout)
{
    return pair_better_il(g, cI, cJ, q, bd, brank, rankJ, dout);
}
static __device__ __forceinline__ bool pair_better(const Ghm& g, const ClusterRec& cI, const ClusterRec& cJ, double q, double bd,
                                                   int brank, int rankJ, double& dout)
{
    return CP ? pair_better_ni(g, cI, cJ, q, bd, brank, rankJ, dout) : pair_better_il(g, cI, cJ, q, bd, brank, rankJ, dout);
}

// best partner of the cluster at position `pos` among the later positions (one warp)
static __device__ void join_row_il(Ghm& g, int o, int pos, int gap, double q)
{
    const int lane = VTID & 31;
    const int n = g.nbc[o];
    const int idI = g.bc[o][pos];
    const ClusterRec& cI = g.cl[idI];
    double bd = gap + 1;
    int bj = -1, brank = 0x7fffffff;
    for (int j = pos + 1 + lane; j < n; j += 32) {
        const int idJ = g.bc[o][j];
        double d;
        if (pair_better(g, cI, g.cl[idJ], q, bd, brank, g.rank[idJ], d)) {
            if (d < gap + 1) {
                bd = d;
                bj = idJ;
                brank = g.rank[idJ];
            }
        }
    }
#pragma unroll
    for (int sft = 16; sft; sft >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, bd, sft);
        const int oj = __shfl_xor_sync(0xffffffffu, bj, sft);
        const int orank = __shfl_xor_sync(0xffffffffu, brank, sft);
        if (oj >= 0 && (bj < 0 || od < bd || (od == bd && orank < brank))) {
            bd = od;
            bj = oj;
            brank = orank;
        }
    }
    if (lane == 0) {
        g.rb_d[idI] = bd;
        g.rb_j[idI] = bj;
    }
}
static __device__ __noinline__ void join_row_ni(Ghm& g, int o, int pos, int gap, double q) { join_row_il(g, o, pos, gap, q); }
static __device__ __forceinline__ void join_row(Ghm& g, int o, int pos, int gap, double q)
{
    if (CP) join_row_ni(g, o, pos, gap, q);
    else join_row_il(g, o, pos, gap, q);
}

// GHM::joinClusters (GHM.cpp:392-404).  The reference rescans all pairs after every merge
// (O(n^3)); here every cluster caches its best later partner and only the rows touched by a
// merge are recomputed.  Vector order is kept as a static rank per pass: erasing a merged
// cluster never changes the relative order of the others.
static __device__ void join_pass(Ghm& g, int gap, int o, double q, Red* red, int* sh_s, long long* pf2)
{
#define JP(k) do { const long long t_ = clock64(); if (VTID == 0) pf2[k] += t_ - jt; jt = t_; } while (0)
    long long jt = clock64();
    const int tid = VTID, warp = tid >> 5;
    int n = g.nbc[o];
    if (n < 2) return;
    for (int i = tid; i < n; i += NT) {
        g.rank[g.bc[o][i]] = i;
        g.pos[g.bc[o][i]] = i;
    }
    __syncthreads();
    for (int row = warp; row < n - 1; row += NW) join_row(g, o, row, gap, q);
    if (tid == 0) {
        g.rb_d[g.bc[o][n - 1]] = gap + 1;
        g.rb_j[g.bc[o][n - 1]] = -1;
    }
    __syncthreads();
    JP(0);
    while (true) {
        // global minimum over the rows: (distance, rank of i, rank of j)
        double bd = gap + 1;
        long long bk = -1;
        for (int i = tid; i < n; i += NT) {
            const int id = g.bc[o][i];
            const int j = g.rb_j[id];
            if (j < 0) continue;
            const double d = g.rb_d[id];
            // positions are monotone in rank: the same lexicographic order, and no search afterwards
            const long long k = ((long long)i << 32) | (unsigned)g.pos[j];
            if (bk < 0 || d < bd || (d == bd && k < bk)) {
                bd = d;
                bk = k;
            }
        }
        block_min_pair(bd, bk, red);
        JP(1);
        if (bk < 0) break;
        if (VTID == 0) pf2[7]++;
        const int pA = (int)(bk >> 32), pB = (int)(bk & 0xffffffff);
        const int idA = g.bc[o][pA], idB = g.bc[o][pB];
        const int rankA = g.rank[idA];
        __syncthreads();
        JP(2);
        if (!merge_coop(g, idA, idB)) return;
        JP(3);
        if (tid == 0) {
            g.cl[idB].dead = 1;
            if (g.tight) {   // the slot can hold a later seed of this orientation
                g.cl[idB].mnext = g.free_head[o];
                g.free_head[o] = idB;
            }
            sh_s[2] = 0;  // rows to recompute
        }
        __syncthreads();
        // erase position pB from the vector (order preserved)
        {
            int v = 0;
            for (int basei = pB; basei < n - 1; basei += NT) {
                const int i = basei + tid;
                if (i < n - 1) v = g.bc[o][i + 1];
                __syncthreads();
                if (i < n - 1) {
                    g.bc[o][i] = v;
                    g.pos[v] = i;
                }
                __syncthreads();
            }
            n--;
            if (tid == 0) g.nbc[o] = n;
            __syncthreads();
        }
        JP(4);
        // refresh the cache: rows that pointed at A or B (and A's own row) are recomputed in full,
        // earlier rows only re-test their pair with the merged cluster
        const ClusterRec& cA = g.cl[idA];
        for (int i = tid; i < n; i += NT) {
            const int id = g.bc[o][i];
            const int j = g.rb_j[id];
            if (id == idA || j == idA || j == idB) {
                if (i < n - 1) g.redo[atomicAdd(&sh_s[2], 1)] = i;
                else {
                    g.rb_d[id] = gap + 1;
                    g.rb_j[id] = -1;
                }
            } else if (i < pA) {
                double d;
                const double cur = (j >= 0) ? g.rb_d[id] : (double)(gap + 1);
                const int crank = (j >= 0) ? g.rank[j] : 0x7fffffff;
                if (pair_better(g, g.cl[id], cA, q, cur, crank, rankA, d) && d < gap + 1) {
                    g.rb_d[id] = d;
                    g.rb_j[id] = idA;
                }
            }
        }
        __syncthreads();
        JP(5);
        const int nredo = sh_s[2];
        for (int r = warp; r < nredo; r += NW) join_row(g, o, g.redo[r], gap, q);
        __syncthreads();
        JP(6);
    }
#undef JP
}

// ---- two-phase join (EXPERIMENTAL, IADH_JOIN2=1, single-warp CTAs of the pass-per-launch path) ----
// Same decisions as join_pass.  What changes is how the warp spends its lanes (of
// the row chunks join_row walks, a third run the expensive pair evaluation for a single useful lane):
//   scan     all lanes test pairs with the integer hull boxes only (exact rejections, see j2_box) and
//            append the survivors, in row-major order, to a 64-entry queue in shared memory;
//   evaluate whenever 32 survivors are queued every lane evaluates one of them (distance, merged r^2,
//            interval test), a segmented minimum over equal rows folds them into rb_d / rb_j.
// A row's result is the lexicographic minimum (distance, rank of the partner) over its valid pairs with
// distance < gap + 1 -- the definition join_row implements -- whichever lane evaluates which pair.
// The boxes live in arrays that are idle during the join: xr (x range), act (y low), act_next (y high).
static __device__ __forceinline__ void j2_box(Ghm& g, int pos, const ClusterRec& c)
{
    // the extents pair_better tests before the distance, rounded outwards; non-finite or huge values
    // (degenerate regressions) become "never apart"
    const uint32_t s = g.pool[c.off], e = g.pool[c.off + c.n - 1];
    const double ylo = cm::mind(cm::mind(c.y_end1, c.y_end2), (double)min(KEY_Y(s), KEY_Y(e)));
    const double yhi = cm::maxd(cm::maxd(c.y_end1, c.y_end2), (double)max(KEY_Y(s), KEY_Y(e)));
    const double lim = 1e9;
    int2 x;
    x.x = (c.x_end1 > -lim && c.x_end1 < lim) ? (int)floor(c.x_end1) : -(1 << 30);
    x.y = (c.x_end2 > -lim && c.x_end2 < lim) ? (int)ceil(c.x_end2) : (1 << 30);
    g.xr[pos] = x;
    g.act[pos] = (ylo > -lim && ylo < lim) ? (int)floor(ylo) : -(1 << 30);
    g.act_next[pos] = (yhi > -lim && yhi < lim) ? (int)ceil(yhi) : (1 << 30);
}
// hulls more than cap = gap + 1 apart in x or y: pair_better rejects such a pair (its bd <= gap + 1)
static __device__ __forceinline__ bool j2_apart(const Ghm& g, int a, int b, int cap)
{
    const int2 ax = g.xr[a], bx = g.xr[b];
    return (bx.x - ax.y > cap) | (ax.x - bx.y > cap) | (g.act[b] - g.act_next[a] > cap) | (g.act[a] - g.act_next[b] > cap);
}

// lanes 0..cnt-1 evaluate queue entries (row position, partner position); runs of equal rows are contiguous
static __device__ __noinline__ void j2_drain(Ghm& g, int o, int gap, double q, const int* qrow, const int* qj, int cnt)
{
    const unsigned full = 0xffffffffu;
    const int lane = VTID & 31;
    int row = -1 - lane, idJ = -1, rk = 0x7fffffff;
    double d = 1e300;
    if (lane < cnt) {
        row = qrow[lane];
        const int id = g.bc[o][qj[lane]];
        const ClusterRec& cI = g.cl[g.bc[o][row]];
        const ClusterRec& cJ = g.cl[id];
        const uint32_t s1 = g.pool[cI.off], e1 = g.pool[cI.off + cI.n - 1];
        const uint32_t s2 = g.pool[cJ.off], e2 = g.pool[cJ.off + cJ.n - 1];
        const double dd = cm::dist_to_cluster(cI, cJ, s1, e1, s2, e2);
        if (dd < gap + 1 && cm::r2_merged(cI, cJ) >= q && cm::partial_overlap(cI, cJ, g.pool)) {
            d = dd;
            idJ = id;
            rk = g.rank[id];
        }
    }
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double od = __shfl_down_sync(full, d, off);
        const int orow = __shfl_down_sync(full, row, off);
        const int oid = __shfl_down_sync(full, idJ, off);
        const int ork = __shfl_down_sync(full, rk, off);
        if (lane + off < 32 && orow == row && oid >= 0 && (idJ < 0 || od < d || (od == d && ork < rk))) {
            d = od;
            idJ = oid;
            rk = ork;
        }
    }
    const int prow = __shfl_up_sync(full, row, 1);
    if (lane < cnt && (lane == 0 || prow != row) && idJ >= 0) {
        const int idI = g.bc[o][row];
        const int cj = g.rb_j[idI];
        if (cj < 0 || d < g.rb_d[idI] || (d == g.rb_d[idI] && rk < g.rank[cj])) {
            g.rb_d[idI] = d;
            g.rb_j[idI] = idJ;
        }
    }
    __syncwarp();
}

static __device__ void join_pass2(Ghm& g, int gap, int o, double q, Red* red, int* sh_s, long long* pf2)
{
    static_assert(NT == 32, "join_pass2 is written for single-warp CTAs");
    __shared__ int qrow[64], qj[64];
    const unsigned full = 0xffffffffu;
    const int tid = VTID, lane = tid & 31;
    int n = g.nbc[o];
    if (n < 2) return;
    const int cap = gap + 1;
    int qn = 0;  // queued survivors (warp-uniform)
    auto push = [&](bool pass, int row, int j) {
        const unsigned m = __ballot_sync(full, pass);
        if (m == 0) return;
        if (pass) {
            const int k = qn + __popc(m & ((1u << lane) - 1u));
            qrow[k] = row;
            qj[k] = j;
        }
        qn += __popc(m);
        __syncwarp();
        if (qn >= 32) {
            j2_drain(g, o, gap, q, qrow, qj, 32);
            const int rest = qn - 32;
            int r0 = 0, j0 = 0;
            if (lane < rest) {
                r0 = qrow[32 + lane];
                j0 = qj[32 + lane];
            }
            __syncwarp();
            if (lane < rest) {
                qrow[lane] = r0;
                qj[lane] = j0;
            }
            qn = rest;
            __syncwarp();
        }
    };
    auto flush = [&]() {
        if (qn) j2_drain(g, o, gap, q, qrow, qj, qn);
        qn = 0;
    };
    for (int i = lane; i < n; i += 32) {
        const int id = g.bc[o][i];
        g.rank[id] = i;
        g.pos[id] = i;
        g.rb_d[id] = gap + 1;
        g.rb_j[id] = -1;
        j2_box(g, i, g.cl[id]);
    }
    __syncthreads();
    {
        // every pair (row < j), flattened: lane k of step t takes pair index 32 t + k in row-major order
        const long long total = (long long)n * (n - 1) / 2;
        int row = 0, j = 1 + lane;
        while (row < n - 1 && j >= n) {
            j = j - n + row + 2;
            row++;
        }
        for (long long base = 0; base < total; base += 32) {
            const bool in = row < n - 1;
            push(in && !j2_apart(g, row, j, cap), row, j);
            j += 32;
            while (row < n - 1 && j >= n) {
                j = j - n + row + 2;
                row++;
            }
        }
        flush();
    }
    __syncthreads();
    while (true) {
        // global minimum over the rows: (distance, position of i, position of j)
        double bd = gap + 1;
        long long bk = -1;
        for (int i = tid; i < n; i += NT) {
            const int id = g.bc[o][i];
            const int j = g.rb_j[id];
            if (j < 0) continue;
            const double d = g.rb_d[id];
            const long long k = ((long long)i << 32) | (unsigned)g.pos[j];
            if (bk < 0 || d < bd || (d == bd && k < bk)) {
                bd = d;
                bk = k;
            }
        }
        block_min_pair(bd, bk, red);
        if (bk < 0) break;
        if (VTID == 0) pf2[7]++;
        const int pA = (int)(bk >> 32), pB = (int)(bk & 0xffffffff);
        const int idA = g.bc[o][pA], idB = g.bc[o][pB];
        __syncthreads();
        if (!merge_coop(g, idA, idB)) return;
        if (tid == 0) {
            g.cl[idB].dead = 1;
            if (g.tight) {   // the slot can hold a later seed of this orientation
                g.cl[idB].mnext = g.free_head[o];
                g.free_head[o] = idB;
            }
            sh_s[2] = 0;  // rows to recompute
        }
        __syncthreads();
        // erase position pB from the vector (order preserved); the boxes move with their clusters
        {
            int v = 0, ylo = 0, yhi = 0;
            int2 bx = make_int2(0, 0);
            for (int basei = pB; basei < n - 1; basei += NT) {
                const int i = basei + tid;
                if (i < n - 1) {
                    v = g.bc[o][i + 1];
                    bx = g.xr[i + 1];
                    ylo = g.act[i + 1];
                    yhi = g.act_next[i + 1];
                }
                __syncthreads();
                if (i < n - 1) {
                    g.bc[o][i] = v;
                    g.pos[v] = i;
                    g.xr[i] = bx;
                    g.act[i] = ylo;
                    g.act_next[i] = yhi;
                }
                __syncthreads();
            }
            n--;
            if (tid == 0) {
                g.nbc[o] = n;
                j2_box(g, pA, g.cl[idA]);
            }
            __syncthreads();
        }
        // rows that pointed at A or B, and A's own row, are recomputed in full (listed in redo);
        // the other earlier rows only get the pair with the merged cluster as a new candidate
        for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            bool cand = false;
            if (i < n) {
                const int id = g.bc[o][i];
                const int j = g.rb_j[id];
                if (id == idA || j == idA || j == idB) {
                    g.rb_d[id] = gap + 1;
                    g.rb_j[id] = -1;
                    if (i < n - 1) g.redo[atomicAdd(&sh_s[2], 1)] = i;
                } else if (i < pA)
                    cand = !j2_apart(g, i, pA, cap);
            }
            push(cand, i, pA);
        }
        __syncthreads();
        const int nredo = sh_s[2];
        for (int r = 0; r < nredo; r++) {
            const int row = g.redo[r];
            for (int base = row + 1; base < n; base += 32) {
                const int j = base + lane;
                push(j < n && !j2_apart(g, row, j, cap), row, j);
            }
        }
        flush();
        __syncthreads();
    }
}

// ---- probability filter ---------------------------------------------------------------------

// BaseCluster::calculateProbability (BaseCluster.cpp:548-590)
static __device__ void calc_probability(Ghm& g, ClusterRec& c)
{
    const double density = (double)g.cnt[c.orient] / g.area;
    double probability = 1.0;
    int px = 0, py = 0;
    bool have = false;
    for (int i = 0; i < c.n; i++) {
        if (!g.pflag[c.off + i]) continue;
        const uint32_t k = g.pool[c.off + i];
        const int x = KEY_X(k), y = KEY_Y(k);
        if (have) {
            const double distance = cm::dpd_d(px, py, x, y);
            const int cc = (int)ceil(0.5 * distance * distance);
            const double p_i = 1.0 - cm::pow_ref(1.0 - density, cc);
            probability *= p_i;
        }
        have = true;
        px = x;
        py = y;
    }
    c.p = probability;
    // duplicates of the first AP are erased without touching the statistics
    const uint32_t f = g.pool[c.off];
    int w = 1;
    for (int i = 1; i < c.n; i++) {
        const uint32_t k = g.pool[c.off + i];
        if (k == f) continue;
        g.pool[c.off + w] = k;
        g.pflag[c.off + w] = g.pflag[c.off + i];
        w++;
    }
    if (w != c.n) {
        c.n = w;
        cm::recompute_sums(c, g.pool);
    }
}

// libstdc++ std::sort (introsort + final insertion sort) restated on cluster ids ordered by p,
// so that ties come out in exactly the reference's order (GHM.cpp:527).
struct PLess {
    const ClusterRec* cl;
    __device__ bool operator()(int l, int r) const { return cl[l].p < cl[r].p; }
};
static __device__ void adjust_heap(int* first, int hole, int len, int value, PLess comp)
{
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (comp(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    int parent = (hole - 1) / 2;
    while (hole > top && comp(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
static __device__ void heap_sort(int* first, int len, PLess comp)
{
    if (len >= 2) {
        int parent = (len - 2) / 2;
        while (true) {
            const int value = first[parent];
            adjust_heap(first, parent, len, value, comp);
            if (parent == 0) break;
            parent--;
        }
    }
    int last = len;
    while (last > 1) {
        --last;
        const int value = first[last];
        first[last] = first[0];
        adjust_heap(first, 0, last, value, comp);
    }
}
static __device__ void unguarded_linear_insert(int* last, PLess comp)
{
    const int val = *last;
    int* next = last - 1;
    while (comp(val, *next)) {
        *last = *next;
        last = next;
        --next;
    }
    *last = val;
}
static __device__ void insertion_sort(int* first, int* last, PLess comp)
{
    if (first == last) return;
    for (int* i = first + 1; i != last; ++i) {
        if (comp(*i, *first)) {
            const int val = *i;
            for (int* p = i; p != first; --p) *p = *(p - 1);
            *first = val;
        } else
            unguarded_linear_insert(i, comp);
    }
}
static __device__ void std_sort(int* first, int n, PLess comp)
{
    if (n <= 0) return;
    // __introsort_loop, recursion on the right part unrolled with an explicit stack
    int depth0 = 0;
    for (int t = n; t > 1; t >>= 1) depth0++;
    depth0 *= 2;
    struct Frame {
        int lo, hi, depth;
    } stack[64];
    int sp = 0;
    stack[sp++] = {0, n, depth0};
    while (sp > 0) {
        Frame f = stack[--sp];
        int lo = f.lo, hi = f.hi, depth = f.depth;
        while (hi - lo > 16) {
            if (depth == 0) {
                heap_sort(first + lo, hi - lo, comp);
                break;
            }
            --depth;
            // __move_median_to_first(lo, lo+1, mid, hi-1)
            int* a = first + lo + 1;
            int* b = first + lo + (hi - lo) / 2;
            int* c = first + hi - 1;
            int* r = first + lo;
            int t;
            if (comp(*a, *b)) {
                if (comp(*b, *c)) { t = *r; *r = *b; *b = t; }
                else if (comp(*a, *c)) { t = *r; *r = *c; *c = t; }
                else { t = *r; *r = *a; *a = t; }
            } else if (comp(*a, *c)) { t = *r; *r = *a; *a = t; }
            else if (comp(*b, *c)) { t = *r; *r = *c; *c = t; }
            else { t = *r; *r = *b; *b = t; }
            // __unguarded_partition(lo+1, hi, pivot = lo)
            int* pf = first + lo + 1;
            int* pl = first + hi;
            const int* pivot = first + lo;
            while (true) {
                while (comp(*pf, *pivot)) ++pf;
                --pl;
                while (comp(*pivot, *pl)) --pl;
                if (!(pf < pl)) break;
                t = *pf; *pf = *pl; *pl = t;
                ++pf;
            }
            const int cut = (int)(pf - first);
            // recurse on [cut, hi) first (the reference does), then loop on [lo, cut)
            // order of processing does not change the result: the two ranges are disjoint
            if (sp < 64) stack[sp++] = {cut, hi, depth};
            hi = cut;
        }
    }
    // __final_insertion_sort
    if (n > 16) {
        insertion_sort(first, first + 16, comp);
        for (int* i = first + 16; i != first + n; ++i) unguarded_linear_insert(i, comp);
    } else
        insertion_sort(first, first + n, comp);
}

// bitonic sort of g.tmp[0..m) (m power of two) by the whole CTA
static __device__ void block_bitonic(uint32_t* a, int m)
{
    for (int k = 2; k <= m; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = VTID; i < m; i += NT) {
                const int l = i ^ j;
                if (l > i) {
                    const uint32_t x = a[i], y = a[l];
                    const bool up = ((i & k) == 0);
                    if ((x > y) == up) {
                        a[i] = y;
                        a[l] = x;
                    }
                }
            }
            __syncthreads();
        }
}

// the same network with every comparator ascending (first step of a merge block: partner i ^ (k - 1)) sorts
// any m: elements at virtual positions >= m are +inf and never move, so their comparators are skipped
static __device__ void block_sort_any(uint32_t* a, int m)
{
    for (int k = 2; (k >> 1) < m; k <<= 1) {
        for (int i = VTID; i < m; i += NT) {
            const int l = i ^ (k - 1);
            if (l > i && l < m) {
                const uint32_t x = a[i], y = a[l];
                if (x > y) {
                    a[i] = y;
                    a[l] = x;
                }
            }
        }
        __syncthreads();
        for (int j = k >> 2; j > 0; j >>= 1) {
            for (int i = VTID; i < m; i += NT) {
                const int l = i ^ j;
                if (l > i && l < m) {
                    const uint32_t x = a[i], y = a[l];
                    if (x > y) {
                        a[i] = y;
                        a[l] = x;
                    }
                }
            }
            __syncthreads();
        }
    }
}

// BaseCluster::calculateProbability by one warp without scratch arrays (resident view): 32 anchor points at
// a time; a backbone point's predecessor is the nearest flagged lane below it (or the last backbone point of
// the previous chunk), the factors are multiplied by lane 0 in backbone order -- the reference's running product.
static __device__ void calc_probability_regs(Ghm& g, ClusterRec& c)
{
    const unsigned full = 0xffffffffu;
    const int lane = VTID & 31;
    const double density = (double)g.cnt[c.orient] / g.area;
    const uint32_t* ap = g.pool + c.off;
    const uint8_t* fl = g.pflag + c.off;
    double probability = 1.0;
    uint32_t prev = 0;
    bool have = false;
    for (int i0 = 0; i0 < c.n; i0 += 32) {
        const int i = i0 + lane;
        const bool f = (i < c.n) && fl[i];
        const uint32_t k = f ? ap[i] : 0u;
        const unsigned bal = __ballot_sync(full, f);
        if (bal == 0) continue;
        const unsigned below = bal & ((1u << lane) - 1u);
        const int src = below ? 31 - __clz(below) : -1;
        const uint32_t pk = __shfl_sync(full, k, src < 0 ? 0 : src);
        const uint32_t a = src >= 0 ? pk : prev;
        double fac = 1.0;
        if (f && (src >= 0 || have)) {
            const double distance = cm::dpd_d(KEY_X(a), KEY_Y(a), KEY_X(k), KEY_Y(k));
            const int cc = (int)ceil(0.5 * distance * distance);
            fac = 1.0 - cm::pow_ref(1.0 - density, cc);
        }
        unsigned m = bal;
        if (!have) m &= m - 1;   // the first backbone point of the cluster has no factor
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const double v = __shfl_sync(full, fac, b);
            if (lane == 0) probability *= v;
        }
        prev = __shfl_sync(full, k, 31 - __clz(bal));
        have = true;
    }
    if (lane == 0) {
        c.p = probability;
        // duplicates of the first AP are erased without touching the statistics
        const uint32_t f = g.pool[c.off];
        int w = 1;
        for (int i = 1; i < c.n; i++) {
            const uint32_t k = g.pool[c.off + i];
            if (k == f) continue;
            if (w != i) {
                g.pool[c.off + w] = k;
                g.pflag[c.off + w] = g.pflag[c.off + i];
            }
            w++;
        }
        if (w != c.n) {
            c.n = w;
            cm::recompute_sums(c, g.pool);
        }
    }
    __syncwarp();
}

// BaseCluster::calculateProbability by one warp: the factors 1 - pow(1 - density, ceil(dpd^2 / 2)) of
// consecutive backbone points are independent (each a ~400-cycle libm-exact pow), only their product
// is ordered.  Lanes compute the factors into `fac`, lane 0 multiplies them in backbone order --
// the same sequence of roundings as the reference's running product.  `bb` / `fac` = this cluster's
// scratch slices (c.n entries each).
static __device__ void calc_probability_warp(Ghm& g, ClusterRec& c, uint32_t* bb, double* fac)
{
    const int lane = VTID & 31;
    const double density = (double)g.cnt[c.orient] / g.area;
    const uint32_t* ap = g.pool + c.off;
    const uint8_t* fl = g.pflag + c.off;
    int m = 0;  // backbone points, in multiset order
    for (int i0 = 0; i0 < c.n; i0 += 32) {
        const int i = i0 + lane;
        const bool f = (i < c.n) && fl[i];
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (f) bb[m + __popc(bal & ((1u << lane) - 1u))] = ap[i];
        m += __popc(bal);
    }
    __syncwarp();
    for (int j = 1 + lane; j < m; j += 32) {
        const uint32_t a = bb[j - 1], b = bb[j];
        const double distance = cm::dpd_d(KEY_X(a), KEY_Y(a), KEY_X(b), KEY_Y(b));
        const int cc = (int)ceil(0.5 * distance * distance);
        fac[j] = 1.0 - cm::pow_ref(1.0 - density, cc);
    }
    __syncwarp();
    if (lane == 0) {
        double probability = 1.0;
        for (int j = 1; j < m; j++) probability *= fac[j];
        c.p = probability;
        // duplicates of the first AP are erased without touching the statistics
        const uint32_t f = g.pool[c.off];
        int w = 1;
        for (int i = 1; i < c.n; i++) {
            const uint32_t k = g.pool[c.off + i];
            if (k == f) continue;
            if (w != i) {
                g.pool[c.off + w] = k;
                g.pflag[c.off + w] = g.pflag[c.off + i];
            }
            w++;
        }
        if (w != c.n) {
            c.n = w;
            cm::recompute_sums(c, g.pool);
        }
    }
    __syncwarp();
}

// GHM::filterBaseClusters / BF / FDR (GHM.cpp:446-540)
static __device__ void filter_pass(Ghm& g, const DevParams& prm, int* sh_s)
{
    const int tid = VTID, warp = tid >> 5;
    // scratch slice of every cluster: exclusive prefix of the cluster sizes (staged in `redo`)
    if (tid == 0) {
        int run = 0, k = 0;
        for (int o = 0; o < 2; o++)
            for (int i = 0; i < g.nbc[o]; i++) {
                g.redo[k++] = run;
                run += g.cl[g.bc[o][i]].n;
            }
        sh_s[5] = run;
    }
    __syncthreads();
    if (sh_s[5] <= g.scratch_cap) {
        const int n0 = g.nbc[0], nall = g.nbc[0] + g.nbc[1];
        for (int k = warp; k < nall; k += NW) {
            ClusterRec& c = g.cl[k < n0 ? g.bc[0][k] : g.bc[1][k - n0]];
            calc_probability_warp(g, c, g.tmp + g.redo[k], g.dscr + g.redo[k]);
        }
    } else if (g.tight) {  // resident view: no scratch slices, one warp per cluster on registers
        const int n0 = g.nbc[0], nall = g.nbc[0] + g.nbc[1];
        for (int k = warp; k < nall; k += NW) calc_probability_regs(g, g.cl[k < n0 ? g.bc[0][k] : g.bc[1][k - n0]]);
    } else {  // (duplicated dots beyond the scratch capacity: one thread per cluster)
        for (int o = 0; o < 2; o++)
            for (int i = tid; i < g.nbc[o]; i += NT) calc_probability(g, g.cl[g.bc[o][i]]);
    }
    __syncthreads();
    for (int o = 0; o < 2; o++) {
        if (prm.mult_hyp == 2) {
            if (tid == 0) {
                PLess comp{g.cl};
                std_sort(g.bc[o], g.nbc[o], comp);
                int k;
                for (k = 0; k < g.nbc[o]; k++)
                    if (g.cl[g.bc[o][k]].p > (k + 1) * prm.prob_cutoff / g.cnt[o]) break;
                for (int i = k; i < g.nbc[o]; i++) g.cl[g.bc[o][i]].dead = 1;
                g.nbc[o] = k;  // filtered APs are NOT returned to the matrix
            }
            __syncthreads();
            continue;
        }
        // none / Bonferroni: failing clusters give their APs back to matrix[o]
        if (tid == 0) {
            const double thr = (prm.mult_hyp == 1) ? (prm.prob_cutoff / g.cnt[o]) : prm.prob_cutoff;
            int w = 0, removed = 0;
            for (int i = 0; i < g.nbc[o]; i++) {
                const int id = g.bc[o][i];
                if (g.cl[id].p > thr) {
                    g.cl[id].dead = 2;  // to be re-inserted
                    removed++;
                } else
                    g.bc[o][w++] = id;
            }
            g.nbc[o] = w;
            sh_s[0] = removed;
            sh_s[1] = 0;
            sh_s[6] = 0;
        }
        __syncthreads();
        if (sh_s[0] == 0) continue;
        if (g.tight) {
            // resident view: the staging area and the matrix are budgets -- count first (alive dots + the
            // anchor points coming back), give up (status 3: host re-runs the GHM on the global path) if over
            int cntl = 0;
            for (int i = tid; i < g.nk[o]; i += NT) cntl += g.alive[o][i] ? 1 : 0;
            for (int c = g.cl_lo[o] + tid; c < g.ncl[o]; c += NT)
                if (g.cl[c].dead == 2) cntl += g.cl[c].n;
            if (cntl) atomicAdd(&sh_s[6], cntl);
            __syncthreads();
            if (sh_s[6] > g.tmp_cap || sh_s[6] > g.key_cap[o]) {
                if (tid == 0) g.status = 3;
                __syncthreads();
                return;
            }
        }
        // gather: alive dots of matrix[o] + APs of the removed clusters, then sort + unique
        {
            const uint32_t* key = g.key[o];
            const uint8_t* alive = g.alive[o];
            const int n = g.nk[o];
            for (int i = tid; i < n; i += NT)
                if (alive[i]) g.tmp[atomicAdd(&sh_s[1], 1)] = key[i];
            // (clusters of orientation o are exactly the ids of region o)
            for (int c = g.cl_lo[o]; c < g.ncl[o]; c++) {
                const ClusterRec& cr = g.cl[c];
                if (cr.dead != 2) continue;
                for (int i = tid; i < cr.n; i += NT) g.tmp[atomicAdd(&sh_s[1], 1)] = g.pool[cr.off + i];
            }
            __syncthreads();
            for (int c = g.cl_lo[o] + tid; c < g.ncl[o]; c += NT)
                if (g.cl[c].dead == 2) g.cl[c].dead = 1;
            const int m = sh_s[1];
            if (g.tight) {
                __syncthreads();
                block_sort_any(g.tmp, m);
            } else {
                int mp = 1;
                while (mp < m) mp <<= 1;
                for (int i = m + tid; i < mp; i += NT) g.tmp[i] = 0xffffffffu;
                __syncthreads();
                block_bitonic(g.tmp, mp);
            }
            // unique-compact back into key[o] (single pass by thread 0 is O(m); use a parallel
            // flag + scan-free approach: position = number of distinct predecessors)
            if (tid == 0) sh_s[1] = 0;
            __syncthreads();
            // chunked ordered compaction
            for (int base = 0; base < m; base += NT) {
                const int i = base + tid;
                const bool keep = (i < m) && (i == 0 || g.tmp[i] != g.tmp[i - 1]);
                // block exclusive scan of keep via ballot
                const unsigned bal = __ballot_sync(0xffffffffu, keep);
                const int lane = tid & 31, w = tid >> 5;
                __shared__ int wsum[NW];
                if (lane == 0) wsum[w] = __popc(bal);
                __syncthreads();
                int off = sh_s[1];
                for (int ww = 0; ww < w; ww++) off += wsum[ww];
                off += __popc(bal & ((1u << lane) - 1u));
                if (keep) {
                    g.key[o][off] = g.tmp[i];
                    g.alive[o][off] = 1;
                }
                __syncthreads();
                if (tid == 0) {
                    int s = 0;
                    for (int ww = 0; ww < NW; ww++) s += wsum[ww];
                    sh_s[1] += s;
                }
                __syncthreads();
            }
            if (tid == 0) g.nk[o] = sh_s[1];
            __syncthreads();
        }
    }
}

// ---- metaclustering ---------------------------------------------------------------------------

// BaseCluster::twistCluster (BaseCluster.cpp:700-719): y -> max_y + min_y - y, statistics
// recomputed.  Short clusters: one thread each; long ones: the whole CTA per cluster (the mirrored
// integer sums follow in closed form: sy' = nM - sy, sxy' = M sx - sxy, syy' = nM^2 - 2M sy + syy).
static __device__ void twist_all(Ghm& g, int o, int* sh_i, int* sh_s)
{
    const int tid = VTID;
    if (tid == 0) sh_s[4] = 0;
    __syncthreads();
    for (int i = tid; i < g.nbc[o]; i += NT) {
        ClusterRec& c = g.cl[g.bc[o][i]];
        if (c.n >= COOP_T(3)) {
            g.redo[atomicAdd(&sh_s[4], 1)] = g.bc[o][i];
            continue;
        }
        int mx = KEY_Y(g.pool[c.off]), mn = mx;
        for (int k = 1; k < c.n; k++) {
            const int y = KEY_Y(g.pool[c.off + k]);
            mx = max(mx, y);
            mn = min(mn, y);
        }
        for (int k = 0; k < c.n; k++) {
            const uint32_t key = g.pool[c.off + k];
            g.pool[c.off + k] = MAKE_KEY(KEY_X(key), mx + mn - KEY_Y(key));
        }
        cm::recompute_sums(c, g.pool);
        cm::update_stats(c, g.pool);
    }
    __syncthreads();
    const int nbig = sh_s[4];
    for (int r = 0; r < nbig; r++) {
        ClusterRec& c = g.cl[g.redo[r]];
        uint32_t* ap = g.pool + c.off;
        const int n = c.n;
        int mn = 0x7fffffff, mx = -1;
        for (int k = tid; k < n; k += NT) {
            const int y = KEY_Y(ap[k]);
            mn = min(mn, y);
            mx = max(mx, y);
        }
        mn = block_min_int(mn, sh_i);
        mx = -block_min_int(-mx, sh_i);
        const int M = mn + mx;  // max_y + min_y
        for (int k = tid; k < n; k += NT) {
            const uint32_t key = ap[k];
            ap[k] = MAKE_KEY(KEY_X(key), M - KEY_Y(key));
        }
        if (tid == 0) {
            const long long Ml = M, nl = n;
            const long long sy = c.sy, sxy = c.sxy, syy = c.syy;
            c.sy = nl * Ml - sy;
            c.sxy = Ml * c.sx - sxy;
            c.syy = nl * Ml * Ml - 2 * Ml * sy + syy;
        }
        __syncthreads();
        stats_coop(g, c);
    }
}

// GHM::clusterClusters (GHM.cpp:542-625)
static __device__ void cluster_clusters(Ghm& g, const DevParams& prm, Red* red, int* sh_i, int* sh_s)
{
    const int tid = VTID;
    bool join_flag = true;
    while (join_flag) {
        join_flag = false;
        for (int k = 0; k < 4; k++) {
            const int ox = (k < 2) ? 1 : 0;
            const int oy = (k % 2 == 0) ? 1 : 0;
            const bool twisted = (ox != oy);
            if (twisted) twist_all(g, oy, sh_i, sh_s);
            int pi = -1, pj = -1;
            const bool found = closest_clusters(g, prm.cluster_gap, ox, oy, prm.q_value, pi, pj, red);
            if (twisted) twist_all(g, oy, sh_i, sh_s);
            if (tid == 0) {
                sh_s[0] = 0;
                if (found) {
                    int cI = g.bc[ox][pi], cJ = g.bc[oy][pj];
                    if (!twisted || !g.cl[cI].was_twisted) {
                        g.cl[cJ].was_twisted = twisted ? 1 : 0;
                        sh_s[0] = 1;
                        const int mi = g.cl[cI].mult, mj = g.cl[cJ].mult;
                        if (mi >= 0 && mj >= 0) {
                            // move every cluster of mj to the end of mi (Multiplicon.cpp:104-111)
                            for (int c = g.mult[mj].head; c >= 0; c = g.cl[c].mnext) g.cl[c].mult = mi;
                            g.cl[g.mult[mi].tail].mnext = g.mult[mj].head;
                            g.mult[mi].tail = g.mult[mj].tail;
                            g.mult[mi].count += g.mult[mj].count;
                            g.mult[mj].head = g.mult[mj].tail = -1;
                            g.mult[mj].count = 0;
                        } else if (mi >= 0 || mj >= 0) {
                            if (mi < 0) {
                                const int t = cI;
                                cI = cJ;
                                cJ = t;
                            }
                            const int m = g.cl[cI].mult;
                            g.cl[cJ].mult = m;
                            g.cl[cJ].mnext = -1;
                            g.cl[g.mult[m].tail].mnext = cJ;
                            g.mult[m].tail = cJ;
                            g.mult[m].count++;
                        } else {
                            const int m = g.nmult++;
                            g.mult[m].head = cI;
                            g.mult[m].tail = cJ;
                            g.mult[m].count = 2;
                            g.mult[m].alive = 1;
                            g.cl[cI].mult = m;
                            g.cl[cI].mnext = cJ;
                            g.cl[cJ].mult = m;
                            g.cl[cJ].mnext = -1;
                        }
                    }
                }
            }
            __syncthreads();
            if (sh_s[0]) join_flag = true;
            __syncthreads();
        }
    }
    if (tid == 0) {
        // singletons (GHM.cpp:600-611), then drop empty / too small multiplicons (613-624)
        for (int o = 0; o < 2; o++)
            for (int i = 0; i < g.nbc[o]; i++) {
                ClusterRec& c = g.cl[g.bc[o][i]];
                if (c.mult < 0) {
                    const int m = g.nmult++;
                    c.was_twisted = 0;
                    c.mult = m;
                    c.mnext = -1;
                    g.mult[m].head = g.mult[m].tail = g.bc[o][i];
                    g.mult[m].count = 1;
                    g.mult[m].alive = 1;
                }
            }
        int kept = 0;
        for (int m = 0; m < g.nmult; m++) {
            unsigned aps = 0;
            for (int c = g.mult[m].head; c >= 0; c = g.cl[c].mnext) aps += (unsigned)g.cl[c].n;
            if (g.mult[m].count == 0 || aps < (unsigned)prm.anchor_points) {
                g.mult[m].alive = 0;
                continue;
            }
            g.out_mults[kept++] = m;
        }
        sh_s[0] = kept;
    }
    __syncthreads();
}

};  // struct KT

// View of one GHM's workspace.  whole = false: the regions of orientation my_o only (phase 1 on its own
// CTA / in its own launch); fresh = true: counters start empty, otherwise they are reloaded from the
// state an earlier CTA or launch left in the task record.
static __device__ void setup_view(Ghm& g, const GhmTask& task, uint8_t* ws, const DevParams& prm, int pool_factor,
                                  int my_o, bool whole, bool fresh)
{
    const int D0 = (int)task.cnt[0], D1 = (int)task.cnt[1], D = D0 + D1;
    const int half_mask = task.half_mask;
    WsLayout L = ws_layout(D, pool_factor);
    uint8_t* base = ws + task.ws_off;
    const int cc0 = D0 / 3 + 2, pc0 = pool_factor * D0 + 32;
    const int eo = whole ? 0 : (my_o ? D0 : 0);       // element (dot) offset of the region
    const int co = whole ? 0 : (my_o ? cc0 : 0);      // cluster-slot offset of the region
    g.key[0] = (uint32_t*)(base + L.key0);
    g.key[1] = (uint32_t*)(base + L.key1);
    g.alive[0] = base + L.alive0;
    g.alive[1] = base + L.alive1;
    g.nxt = (int*)(base + L.nxt) + eo;
    g.pool = (uint32_t*)(base + L.pool);
    g.pflag = base + L.pflag;
    g.cl = (ClusterRec*)(base + L.cl);
    g.cl_lo[0] = 0; g.cl_hi[0] = cc0;
    g.cl_lo[1] = cc0; g.cl_hi[1] = cc0 + D1 / 3 + 2;
    g.pool_lo[0] = 0; g.pool_hi[0] = pc0;
    g.pool_lo[1] = pc0; g.pool_hi[1] = L.pool_cap;
    for (int r = 0; r < 2; r++) {
        const bool have = !fresh && ((half_mask >> r) & 1) && (whole || r == my_o);
        g.nbc[r] = have ? task.half_nbc[r] : 0;
        g.ncl[r] = have ? task.half_ncl[r] : g.cl_lo[r];
        g.pool_used[r] = have ? task.half_pool[r] : g.pool_lo[r];
    }
    if (!whole) {  // the other orientation's regions belong to another CTA
        const int other = my_o ^ 1;
        g.pool_lo[other] = g.pool_hi[other] = g.pool_used[other] = 0;
        g.ncl[other] = g.cl_lo[other];
        g.cl_hi[other] = g.cl_lo[other];
    }
    g.bc[0] = (int*)(base + L.bc0);
    g.bc[1] = (int*)(base + L.bc1);
    g.act = (int*)(base + L.act) + co;
    g.act_next = (int*)(base + L.act_next) + co;
    g.changed = base + L.changed + co;
    g.mult = (MultRec*)(base + L.mult);
    g.nmult = 0;
    g.out_mults = (int*)(base + L.out_mults);
    g.rb_d = (double*)(base + L.rb_d);
    g.rb_j = (int*)(base + L.rb_j);
    g.rank = (int*)(base + L.rank);
    g.pos = (int*)(base + L.pos);
    g.xr = (int2*)(base + L.xr) + co;
    g.redo = (int*)(base + L.redo) + co;
    g.dscr = (double*)(base + L.dscr) + 2 * (size_t)eo;
    g.tmp = (uint32_t*)(base + L.tmp) + eo;
    g.tmpf = (uint8_t*)((uint32_t*)(base + L.tmp) + L.tmp_cap) + eo;
    g.scratch_cap = whole ? D : 0;
    g.key_cap[0] = g.key_cap[1] = 0x7fffffff;   // (worst-case sized: never the limit)
    g.tmp_cap = 0x7fffffff;
    g.dscr_cap = D > 0 ? D : 1;
    g.tight = 0;
    g.ctr = nullptr;
    g.free_head[0] = g.free_head[1] = -1;
    // successor pointers: orientation 1 follows orientation 0 inside the GHM's array (a view of one
    // orientation already starts at its own region)
    g.nxt_off[0] = 0;
    g.nxt_off[1] = whole ? D0 : 0;
    for (int r = 0; r < 2; r++) {
        const bool have = !fresh && ((half_mask >> r) & 1) && (whole || r == my_o);
        g.seeded[r] = have ? (task.half_seed[r] & 1) : 0;
        g.dirty[r] = have ? ((task.half_seed[r] >> 1) & 1) : 0;
        g.g3[r] = have ? task.half_g3[r] : 0;
    }
    g.status = fresh ? 0 : max(task.half_status[0], task.half_status[1]);
    for (int k = 0; k < 4; k++) g.coop[k] = prm.coop[k];
    g.nk[0] = D0;
    g.nk[1] = D1;
    g.cnt[0] = task.cnt[0];
    g.cnt[1] = task.cnt[1];
    g.area = (double)task.area;
    g.level = task.level;
}

// ---- kernel -----------------------------------------------------------------------------------

template <int NT>
__global__ void __launch_bounds__(NT) cluster_kernel(GhmTask* tasks, GhmProf* profs, const int* __restrict__ order,
                                                     int n_tasks, int* next_task,
                                                     const uint32_t* __restrict__ dots, uint8_t* ws,
                                                     DevParams prm, int pool_factor, int* sm_slot, DatasetDev ds,
                                                     ProfileDev pf, int hl)
{
    using K = KT<NT>;
    constexpr int NW = K::NW;
    __shared__ Ghm g;
    __shared__ pk::PGhm pg;
    __shared__ Red red;
    __shared__ int sh_i[NW];
    __shared__ int sh_s[8];
    __shared__ int sh_task;
    __shared__ long long prof[8];
    __shared__ long long prof2[8];
    if (threadIdx.x == 0) {
        const unsigned smid = read_smid();
        s_rot = (NW > 1) ? 32 * (atomicAdd(&sm_slot[smid & 1023], 1) % NW) : 0;
    }
    __syncthreads();
    const int tid = VTID;
#define PROF_BEGIN() long long pt_ = clock64()
#define PROF_END(k)                              \
    do {                                         \
        const long long pn_ = clock64();         \
        if (tid == 0) prof[k] += pn_ - pt_;      \
        pt_ = pn_;                               \
    } while (0)
    while (true) {
        if (tid == 0) sh_task = atomicAdd(next_task, 1);
        __syncthreads();
        const int ti = sh_task;
        if (ti >= n_tasks) break;
        // queue item = GHM index * 4 + orientation mask: 1 or 2 = that orientation only (the GHM has a
        // second item for the other one), 3 = both orientations one after the other, 0 = phase 2 only
        const int item = order[ti];
        const int my_mask = item & 3;
        const int my_o = (my_mask == 2) ? 1 : 0;
        GhmTask& task = tasks[item >> 2];
        GhmProf& tprof = profs[item >> 2];
        const int D0 = (int)task.cnt[0], D1 = (int)task.cnt[1];
        const int half_mask = task.half_mask;
        const bool run_p1 = my_mask != 0;
        const double q = prm.q_value;
        int gmax = 1;   // largest gap size of the schedule (seed_pass)
        for (int k = 0; k < prm.n_gaps && k < 10; k++) gmax = max(gmax, prm.gaps[k]);
        if (tid < 8) { prof[tid] = 0; prof2[tid] = 0; }
        // view: phase 1 = this CTA's orientation only (its regions of every shared array),
        //       phase 2 = the whole GHM
        auto setup = [&](bool whole, bool fresh) { setup_view(g, task, ws, prm, pool_factor, my_o, whole, fresh); };
        // the dots (written once by K1) become the mutable matrix of one orientation
        auto stage = [&](int o) {
            const uint32_t* src = dots + task.dot_off + (o ? D0 : 0);
            const int n = o ? D1 : D0;
            for (int i = tid; i < n; i += NT) {
                g.key[o][i] = src[i];
                g.alive[o][i] = 1;
            }
        };
        PROF_BEGIN();
        const long long t_begin = clock64();
        // ---- phase 1 (GHM.cpp:167-184): seed / enrich / join at growing gaps, this orientation
        if (run_p1) {
            if (tid == 0) setup(my_mask == 3, true);
            __syncthreads();
            for (int o = 0; o < 2; o++)
                if ((my_mask >> o) & 1) stage(o);
            __syncthreads();
            for (int gi = 0; gi < prm.n_gaps && gi < 10 && prm.gaps[gi] > 0 && !g.status; gi++) {
                const int gap = prm.gaps[gi];
                for (int o = 0; o < 2; o++) {
                    if (!((my_mask >> o) & 1)) continue;
                    if (!K::seed_skippable(g, gap, o)) K::seed_pass(g, gap, gmax, o, q, sh_i, sh_s);
                    __syncthreads();
                    PROF_END(0);
                    if (g.status) break;
                    K::enrich_pass(g, gap, o, o, q, sh_i, sh_s);
                    __syncthreads();
                    PROF_END(1);
                    if (g.status) break;
                    K::join_pass(g, gap, o, q, &red, sh_s, prof2);
                    __syncthreads();
                    PROF_END(2);
                    if (g.status) break;
                }
            }
            __syncthreads();
            if (tid == 0) {
                for (int o = 0; o < 2; o++) {
                    if (!((my_mask >> o) & 1)) continue;
                    task.half_nbc[o] = g.nbc[o];
                    task.half_ncl[o] = g.ncl[o];
                    task.half_pool[o] = g.pool_used[o];
                    task.half_status[o] = g.status;
                    tprof.half_cycles[o] = clock64() - t_begin;
                }
            }
        }
        // ---- hand-over: the CTA that finishes the GHM's last orientation carries on
        bool last = true;
        if (task.half_needed > 1) {
            __threadfence();
            __syncthreads();
            if (tid == 0) sh_s[7] = atomicAdd(&task.half_done, 1);
            __syncthreads();
            last = (sh_s[7] + 1 == task.half_needed);
            __threadfence();  // acquire: also drops this SM's L1 lines (the other CTA wrote the rest)
        }
        if (tid == 0)
            for (int k = 0; k < 8; k++) {
                if (k < 5 && prof[k]) atomicAdd((unsigned long long*)&tprof.prof[k], (unsigned long long)prof[k]);
                if (prof2[k]) atomicAdd((unsigned long long*)&tprof.prof2[k], (unsigned long long)prof2[k]);
            }
        __syncthreads();
        if (!last) continue;
        // ---- phase 2 (GHM.cpp:186-217): cross-orientation enrichment, filter, metaclustering
        if (tid < 8) { prof[tid] = 0; prof2[tid] = 0; }
        if (tid == 0) setup(true, false);
        __syncthreads();
        for (int o = 0; o < 2; o++)
            if (!((half_mask >> o) & 1)) stage(o);
        __syncthreads();
        const long long t_p2 = clock64();
        pt_ = t_p2;
        bool ok = (g.status == 0);
        if (ok) do {
            for (int rep = 0; rep < 2 && ok; rep++) {
                for (int e = 0; e < 4; e++) {
                    // GHM.cpp:186-189: (cluster, matrix) orientations (0,0) (1,1) (1,0) (0,1)
                    const int oc = (e == 1 || e == 2) ? 1 : 0;
                    const int om = (e == 1 || e == 3) ? 1 : 0;
                    K::enrich_pass(g, prm.gap_size, oc, om, q, sh_i, sh_s);
                    __syncthreads();
                    PROF_END(1);
                    if (g.status) { ok = false; break; }
                }
                if (!ok) break;
                if (rep == 0) {
                    K::filter_pass(g, prm, sh_s);
                    __syncthreads();
                    if (g.status) ok = false;
                    if (!ok) break;
                    PROF_END(3);
                }
            }
            if (!ok) break;
            K::cluster_clusters(g, prm, &red, sh_i, sh_s);
            PROF_END(4);
        } while (false);
        __syncthreads();
        if (tid == 0) {
            task.status = ok ? 0 : (g.status ? g.status : 1);
            task.n_mult = ok ? sh_s[0] : 0;
            task.n_clusters = (g.ncl[0] - g.cl_lo[0]) + (g.ncl[1] - g.cl_lo[1]);
            // critical path of this GHM: the slower orientation, then phase 2
            const long long h0 = (half_mask & 1) ? tprof.half_cycles[0] : 0, h1 = (half_mask & 2) ? tprof.half_cycles[1] : 0;
            tprof.prof[5] = (h0 > h1 ? h0 : h1) + (clock64() - t_p2);
            for (int k = 0; k < 5; k++)
                if (prof[k]) atomicAdd((unsigned long long*)&tprof.prof[k], (unsigned long long)prof[k]);
        }
        __syncthreads();
        // K3 measure as the tail of the GHM (bounds, HL duplicate filter and false anchor points, sizes)
        if (ok && task.n_mult > 0) {
            if (hl) pk::measure_ghm<true>(pg, task, ws, ds, pf, pool_factor);
            else pk::measure_ghm<false>(pg, task, ws, ds, pf, pool_factor);
        }
        __syncthreads();
    }
}

// ---- resident kernel: the whole clustering state of a GHM in shared memory -----------------------
// One CTA per GHM from the staged dots to the packed sizes (both orientations one after the other, phase 2,
// K3 measuring pass); the state never travels to HBM between passes.  What leaves the SM at the end is what
// the K3 writer reads: the anchor points and records of the clusters that made it into a multiplicon, copied
// to the same positions of the GHM's global workspace (WsLayout), so the writer is unchanged.  A GHM that
// outgrows a budget of RsLayout reports status != 0 and is re-run on the global path by the host.
#ifdef IADH_HOST_EMU
static uint8_t* rs_smem = nullptr;   // tests/emu: the "dynamic shared memory" of the emulated block
#else
extern __shared__ __align__(16) uint8_t rs_smem[];
#endif

static __device__ void setup_view_resident(Ghm& g, const GhmTask& task, uint8_t* base, const RsLayout& L, const DevParams& prm)
{
    g.key[0] = (uint32_t*)(base + L.key0);
    g.key[1] = (uint32_t*)(base + L.key1);
    g.alive[0] = base + L.alive0;
    g.alive[1] = base + L.alive1;
    g.nxt = (int*)(base + L.nxt);
    g.pool = (uint32_t*)(base + L.pool);
    g.pflag = base + L.pflag;
    g.cl = (ClusterRec*)(base + L.cl);
    g.cl_lo[0] = 0; g.cl_hi[0] = L.cl_cap0;
    g.cl_lo[1] = L.cl_cap0; g.cl_hi[1] = L.cl_cap0 + L.cl_cap1;
    // one pool region (the orientations are clustered one after the other by this CTA)
    g.pool_lo[0] = 0; g.pool_hi[0] = L.pool_cap; g.pool_used[0] = 0;
    g.pool_lo[1] = g.pool_hi[1] = g.pool_used[1] = L.pool_cap;
    for (int r = 0; r < 2; r++) {
        g.nbc[r] = 0;
        g.ncl[r] = g.cl_lo[r];
        g.free_head[r] = -1;
    }
    g.bc[0] = (int*)(base + L.bc0);
    g.bc[1] = (int*)(base + L.bc1);
    g.act = (int*)(base + L.act);
    g.act_next = (int*)(base + L.act_next);
    g.changed = base + L.changed;
    g.mult = (MultRec*)(base + L.mult);
    g.nmult = 0;
    g.out_mults = (int*)(base + L.out_mults);
    g.rb_d = (double*)(base + L.rb_d);
    g.rb_j = (int*)(base + L.rb_j);
    g.rank = (int*)(base + L.rank);
    g.pos = (int*)(base + L.pos);
    g.xr = (int2*)(base + L.xr);
    g.redo = (int*)(base + L.redo);
    g.dscr = (double*)(base + L.dscr);
    g.tmp = (uint32_t*)(base + L.tmp);
    g.tmpf = base + L.tmpf;
    g.nxt_off[0] = 0;
    g.nxt_off[1] = L.cap0;
    for (int r = 0; r < 2; r++) g.seeded[r] = g.dirty[r] = g.g3[r] = 0;
    g.scratch_cap = 0;
    g.key_cap[0] = L.cap0;
    g.key_cap[1] = L.cap1;
    g.tmp_cap = L.tmp_cap;
    g.dscr_cap = L.dscr_cap;
    g.tight = 1;
    g.status = 0;
    for (int k = 0; k < 4; k++) g.coop[k] = prm.coop[k];
    g.nk[0] = (int)task.cnt[0];
    g.nk[1] = (int)task.cnt[1];
    g.cnt[0] = task.cnt[0];
    g.cnt[1] = task.cnt[1];
    g.area = (double)task.area;
    g.level = task.level;
}

template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) resident_kernel(GhmTask* tasks, const int* __restrict__ order, int n_tasks, int* next_task,
                                                            const uint32_t* __restrict__ dots, uint8_t* ws, DevParams prm,
                                                            int pool_factor, int smem_bytes, DatasetDev ds, ProfileDev pf, int hl,
                                                            unsigned long long* dbg)
{
    using K = KT<NT, true>;
    constexpr int NW = K::NW;
    __shared__ Ghm g;
    __shared__ pk::PGhm pg;
    __shared__ Red red;
    __shared__ RsLayout lay;
    __shared__ int sh_i[NW];
    __shared__ int sh_s[8];
    __shared__ int sh_task;
    __shared__ long long prof2[8];
    __shared__ long long rctr[16];
    if (threadIdx.x == 0) s_rot = 0;
    if (threadIdx.x < 16) rctr[threadIdx.x] = 0;
    __syncthreads();
    const int tid = VTID, lane = tid & 31, warp = tid >> 5;
    const double q = prm.q_value;
    int gmax = 1;   // largest gap size of the schedule (seed_pass)
    for (int k = 0; k < prm.n_gaps && k < 10; k++) gmax = max(gmax, prm.gaps[k]);
    long long rt_ = 0;
#define RT(k) do { if (dbg) { const long long n_ = clock64(); if (tid == 0) rctr[k] += n_ - rt_; rt_ = n_; } } while (0)
    while (true) {
        if (tid == 0) sh_task = atomicAdd(next_task, 1);
        __syncthreads();
        const int ti = sh_task;
        if (ti >= n_tasks) break;
        if (dbg) rt_ = clock64();
        GhmTask& task = tasks[order[ti]];
        const int D0 = (int)task.cnt[0], D1 = (int)task.cnt[1];
        if (tid == 0) {
            lay = rs_layout(D0, D1, NT, prm.rs_squeeze);
            setup_view_resident(g, task, rs_smem, lay, prm);
            g.ctr = dbg ? rctr : nullptr;
            if ((int)lay.total > smem_bytes) g.status = 4;   // (the host sizes the classes: never expected)
        }
        __syncthreads();
        bool ok = (g.status == 0);
        if (ok) {
            // the dots (written once by K1) become the two mutable matrices
            const uint32_t* src = dots + task.dot_off;
            for (int i = tid; i < D0; i += NT) {
                g.key[0][i] = src[i];
                g.alive[0][i] = 1;
            }
            for (int i = tid; i < D1; i += NT) {
                g.key[1][i] = src[D0 + i];
                g.alive[1][i] = 1;
            }
            __syncthreads();
            RT(0);
            // ---- phase 1 (GHM.cpp:167-184): seed / enrich / join at growing gaps, per orientation
            for (int gi = 0; gi < prm.n_gaps && gi < 10 && prm.gaps[gi] > 0 && !g.status; gi++) {
                const int gap = prm.gaps[gi];
                for (int o = 0; o < 2; o++) {
                    if (g.nk[o] < 3) continue;   // cannot seed a cluster (an orientation's matrix only shrinks in phase 1)
                    if (!K::seed_skippable(g, gap, o)) K::seed_pass(g, gap, gmax, o, q, sh_i, sh_s);
                    __syncthreads();
                    RT(1);
                    if (g.status) break;
                    if (g.nbc[o] > 0) {
                        K::enrich_pass(g, gap, o, o, q, sh_i, sh_s);
                        __syncthreads();
                        RT(2);
                        if (g.status) break;
                    }
                    if (g.nbc[o] > 1) {
                        if constexpr (NT == 32) K::join_pass2(g, gap, o, q, &red, sh_s, prof2);
                        else K::join_pass(g, gap, o, q, &red, sh_s, prof2);
                        __syncthreads();
                        RT(3);
                        if (g.status) break;
                    }
                }
            }
            __syncthreads();
            ok = (g.status == 0);
        }
        // ---- phase 2 (GHM.cpp:186-217): cross-orientation enrichment, filter, metaclustering
        if (ok) do {
            for (int rep = 0; rep < 2 && ok; rep++) {
                for (int e = 0; e < 4; e++) {
                    const int oc = (e == 1 || e == 2) ? 1 : 0;
                    const int om = (e == 1 || e == 3) ? 1 : 0;
                    if (g.nbc[oc] > 0 && g.nk[om] > 0) K::enrich_pass(g, prm.gap_size, oc, om, q, sh_i, sh_s);
                    __syncthreads();
                    if (g.status) { ok = false; break; }
                }
                RT(4);
                if (!ok) break;
                if (rep == 0) {
                    K::filter_pass(g, prm, sh_s);
                    __syncthreads();
                    RT(5);
                    if (g.status) ok = false;
                }
            }
            if (!ok) break;
            K::cluster_clusters(g, prm, &red, sh_i, sh_s);
            RT(6);
        } while (false);
        __syncthreads();
        const int n_mult = ok ? sh_s[0] : 0;
        __syncthreads();
        if (tid == 0) {
            task.status = ok ? 0 : (g.status ? g.status : 1);
            task.n_mult = n_mult;
            task.n_clusters = 0;
            task.n_aps = 0;
            task.out_bytes = 0;
        }
        // K3 measure on the resident state (bounds, HL duplicate filter and false anchor points, sizes);
        // its per-AP scratch (HL only) goes straight to the global workspace, where the writer reads it
        if (ok && n_mult > 0) {
            const WsLayout GL = ws_layout(D0 + D1, pool_factor);
            uint8_t* gbase = ws + task.ws_off;
            if (tid == 0) {
                pg.pool = g.pool;
                pg.cl = g.cl;
                pg.mult = g.mult;
                pg.out_mults = g.out_mults;
                pg.oc = g.act;
                pg.scratch = (int*)(gbase + GL.nxt);
                pg.n_mult = n_mult;
            }
            __syncthreads();
            if (hl) pk::measure_core<true>(pg, task, ds, pf);
            else pk::measure_core<false>(pg, task, ds, pf);
            __syncthreads();
            // hand-over to the writer: same ids, same pool offsets, global arrays
            {
                uint32_t* gpool = (uint32_t*)(gbase + GL.pool);
                ClusterRec* gcl = (ClusterRec*)(gbase + GL.cl);
                MultRec* gmult = (MultRec*)(gbase + GL.mult);
                int* gom = (int*)(gbase + GL.out_mults);
                int* goc = (int*)(gbase + GL.act);
                const int n_oc = pg.n_oc;
                for (int i = warp; i < n_oc; i += NW) {
                    const int id = g.act[i];
                    const ClusterRec& c = g.cl[id];
                    const uint32_t* ap = g.pool + c.off;
                    for (int k = lane; k < c.n; k += 32) gpool[c.off + k] = ap[k];
                    const uint32_t* sw = reinterpret_cast<const uint32_t*>(&c);
                    uint32_t* dw = reinterpret_cast<uint32_t*>(&gcl[id]);
                    for (int k = lane; k < (int)(sizeof(ClusterRec) / 4); k += 32) dw[k] = sw[k];
                }
                for (int m = tid; m < n_mult; m += NT) {
                    const int mi = g.out_mults[m];
                    gom[m] = mi;
                    gmult[mi] = g.mult[mi];
                }
                for (int i = tid; i < n_oc; i += NT) goc[i] = g.act[i];
            }
        }
        __syncthreads();
        RT(7);
        if (dbg && tid == 0) rctr[11] += prof2[7], prof2[7] = 0, rctr[15] += (ok ? 0 : 1);
    }
    if (dbg) {
        __syncthreads();
        if (tid < 16 && rctr[tid]) atomicAdd(&dbg[tid], (unsigned long long)rctr[tid]);
    }
#undef RT
}

// ---- work distribution of the pass-per-launch kernels ---------------------------------------------------
// Most (GHM, orientation) items have nothing to do in most passes (config 4: half of the GHMs never form a
// cluster; 72 % of the join calls see no cluster).  A visit used to cost a queue atomic plus a chain of
// dependent loads into the 168-byte task record before the item could be dismissed.  Now a CTA takes up to 32
// consecutive queue entries with one atomic; lane l reads entry l's packed state word (4 bytes per item,
// coalesced, L2-resident) and only the entries with work in this pass are handed to the CTA.
//   bits 0-7 smallest gap size with a 3-dot chain (saturated, seed_pass), 8 successors computed, 9 a dot was
//   erased since, 10-11 base clusters (saturated at 3), 12 failed (pool exhausted: the host retries the GHM)
struct Batch {
    int item[32];
    int n;
    int more;
};
static __device__ __forceinline__ uint32_t pack_istate(const Ghm& g, int o)
{
    const uint32_t g3 = (uint32_t)(g.g3[o] < 0 ? 0 : (g.g3[o] > 255 ? 255 : g.g3[o]));
    return g3 | (g.seeded[o] ? 0x100u : 0u) | (g.dirty[o] ? 0x200u : 0u) | ((uint32_t)min(g.nbc[o], 3) << 10) | (g.status ? 0x1000u : 0u);
}
// mode 0 seed, 1 enrich, 2 join (items = GHM * 2 + orientation); 3 = a phase-2 stage (items = GHM * 4)
static __device__ void pop_batch(Batch& b, const int* __restrict__ items, int n_items, int* next, const uint32_t* __restrict__ istate,
                                 int mode, int gap, bool first, int vtid)
{
    if (vtid < 32) {
        // the queue is ordered heaviest item first: one entry at a time at its head (load balance), up to 32
        // towards its tail (light items, mostly dismissed: fewer queue atomics)
        int base = 0, cnt = 1;
        if (vtid == 0) {
            const float f = fminf(1.0f, (float)*(volatile int*)next / (float)(n_items > 0 ? n_items : 1));
            cnt = 1 + (int)(31.0f * f * f * f);
            base = atomicAdd(next, cnt);
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        cnt = __shfl_sync(0xffffffffu, cnt, 0);
        bool act = false;
        int it = -1;
        if (vtid < cnt && base + vtid < n_items) {
            it = items[base + vtid];
            if (mode == 3) {
                const uint32_t st = istate[(it >> 2) * 2] | istate[(it >> 2) * 2 + 1];
                act = (st & 0x1c00u) != 0;   // a cluster in either orientation, or a failure to report
            } else {
                const uint32_t st = istate[it];
                if (first) act = true;
                else if (st & 0x1000u) act = false;
                else if (mode == 0) act = !((st & 0x300u) == 0x100u && gap < (int)(st & 0xffu));
                else act = (int)((st >> 10) & 3u) >= mode;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, act);
        if (act) b.item[__popc(m & ((1u << vtid) - 1u))] = it;
        if (vtid == 0) {
            b.n = __popc(m);
            b.more = base < n_items;
        }
    }
    __syncthreads();
}

// ---- phase 1, one pass per launch ---------------------------------------------------------------
// For large batches (thousands of GHMs) the fused kernel above is throughput-bound by instruction
// fetch: co-resident CTAs sit in different passes of a 430 KB image and per-SM progress saturates
// at about two CTAs (profiles/README.md).  Here every launch runs ONE pass (seed, enrich or join at
// one gap size) over all (GHM, orientation) items, so all CTAs of an SM execute the same few KB of
// code and many more of them can be resident; the per-orientation state (cluster count, pool
// level, status) travels between launches in the task record.  Phase 2 then runs through
// cluster_kernel with "phase 2 only" items.  Decisions and their order are unchanged.
template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) phase1_pass_kernel(GhmTask* tasks, const int* __restrict__ items, int n_items,
                                                         int* next_item, const uint32_t* __restrict__ dots, uint8_t* ws,
                                                         DevParams prm, int pool_factor, int* sm_slot, int pass, int gi,
                                                         unsigned long long* dbg, uint32_t* istate)
{
    using K = KT<NT, true>;
    constexpr int NW = K::NW;
    __shared__ Ghm g;
    __shared__ Red red;
    __shared__ int sh_i[NW];
    __shared__ int sh_s[8];
    __shared__ int sh_task;
    __shared__ long long prof2[8];
    if (threadIdx.x == 0) {
        const unsigned smid = read_smid();
        s_rot = (NW > 1) ? 32 * (atomicAdd(&sm_slot[smid & 1023], 1) % NW) : 0;
    }
    __syncthreads();
    const int tid = VTID;
    const int gap = prm.gaps[gi];
    const double q = prm.q_value;
    int gmax = 1;   // largest gap size of the schedule (seed_pass)
    for (int k = 0; k < prm.n_gaps && k < 10; k++) gmax = max(gmax, prm.gaps[k]);
    const bool first = (pass == 0 && gi == 0);
    __shared__ Batch batch;
    while (true) {
      // 32 queue entries at a time; items without work in this pass (failed earlier, no 3-dot chain at this
      // gap size, no cluster to enrich, fewer than two clusters to join) are dismissed on their state word
      pop_batch(batch, items, n_items, next_item, istate, pass, gap, first, tid);
      if (!batch.more) break;
      const int nb = batch.n;
      for (int bi = 0; bi < nb; bi++) {
        const int item = batch.item[bi];   // GHM index * 2 + orientation
        const int o = item & 1;
        GhmTask& task = tasks[item >> 1];
        if (tid < 8) prof2[tid] = 0;
        if (tid == 0) setup_view(g, task, ws, prm, pool_factor, o, false, first);
        __syncthreads();
        if (first) {
            const int D0 = (int)task.cnt[0];
            const uint32_t* src = dots + task.dot_off + (o ? D0 : 0);
            const int n = (int)task.cnt[o];
            for (int i = tid; i < n; i += NT) {
                g.key[o][i] = src[i];
                g.alive[o][i] = 1;
            }
            __syncthreads();
        }
        if (pass == 0) {
            if (!K::seed_skippable(g, gap, o)) K::seed_pass(g, gap, gmax, o, q, sh_i, sh_s);
        }
        else if (pass == 1) {
            if (g.nbc[o] > 0) K::enrich_pass(g, gap, o, o, q, sh_i, sh_s);
        } else {
            if (g.nbc[o] > 1) K::join_pass(g, gap, o, q, &red, sh_s, prof2);
        }
        __syncthreads();
        if (tid == 0) {
            task.half_nbc[o] = g.nbc[o];
            task.half_ncl[o] = g.ncl[o];
            task.half_pool[o] = g.pool_used[o];
            task.half_status[o] = g.status;
            task.half_g3[o] = g.g3[o];
            task.half_seed[o] = (g.seeded[o] ? 1 : 0) | (g.dirty[o] ? 2 : 0);
            istate[item] = pack_istate(g, o);
            if (dbg && pass == 2)  // IADH_TRACE: join sub-phase cycles of the whole batch
                for (int k = 0; k < 8; k++)
                    if (prof2[k]) atomicAdd(&dbg[k], (unsigned long long)prof2[k]);
        }
        __syncthreads();
      }
      __syncthreads();   // (the batch is refilled by warp 0)
    }
}

// The join pass with the two-phase scan (EXPERIMENTAL, IADH_JOIN2=1): a kernel of its own, so that the
// default kernels above stay exactly as measured.  Same contract as phase1_pass_kernel with pass == 2.
template <int MINB>
__global__ void __launch_bounds__(32, MINB) phase1_join2_kernel(GhmTask* tasks, const int* __restrict__ items, int n_items,
                                                                int* next_item, uint8_t* ws, DevParams prm, int pool_factor,
                                                                int gi, unsigned long long* dbg, uint32_t* istate)
{
    constexpr int NT = 32;   // (VTID)
    using K = KT<NT, true>;
    __shared__ Ghm g;
    __shared__ Red red;
    __shared__ int sh_s[8];
    __shared__ int sh_task;
    __shared__ long long prof2[8];
    if (threadIdx.x == 0) s_rot = 0;
    __syncthreads();
    const int tid = VTID;
    const int gap = prm.gaps[gi];
    const double q = prm.q_value;
    __shared__ Batch batch;
    while (true) {
      pop_batch(batch, items, n_items, next_item, istate, 2, gap, false, tid);
      if (!batch.more) break;
      const int nb = batch.n;
      for (int bi = 0; bi < nb; bi++) {
        const int item = batch.item[bi];   // GHM index * 2 + orientation
        const int o = item & 1;
        GhmTask& task = tasks[item >> 1];
        if (tid < 8) prof2[tid] = 0;
        if (tid == 0) setup_view(g, task, ws, prm, pool_factor, o, false, false);
        __syncthreads();
        if (g.nbc[o] > 1) K::join_pass2(g, gap, o, q, &red, sh_s, prof2);
        __syncthreads();
        if (tid == 0) {
            task.half_nbc[o] = g.nbc[o];
            task.half_ncl[o] = g.ncl[o];
            task.half_pool[o] = g.pool_used[o];
            task.half_status[o] = g.status;
            istate[item] = pack_istate(g, o);
            if (dbg)
                for (int k = 0; k < 8; k++)
                    if (prof2[k]) atomicAdd(&dbg[k], (unsigned long long)prof2[k]);
        }
        __syncthreads();
      }
      __syncthreads();
    }
}

// ---- phase 2, one stage per launch (same idea; whole-GHM view) ---------------------------------
// stage 0: the four cross-orientation enrichments (GHM.cpp:186-189); 1: probability filter
// (GHM.cpp:198-208); 2: the four enrichments again (:210-213); 3: metaclustering (:215-217) + the
// measuring pass of K3.
template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) phase2_stage_kernel(GhmTask* tasks, const int* __restrict__ items, int n_items,
                                                                int* next_item, const uint32_t* __restrict__ dots,
                                                                uint8_t* ws, DevParams prm, int pool_factor, int* sm_slot,
                                                                DatasetDev ds, ProfileDev pf, int hl, int stage,
                                                                const uint32_t* __restrict__ istate)
{
    using K = KT<NT, true>;
    constexpr int NW = K::NW;
    __shared__ Ghm g;
    __shared__ pk::PGhm pg;
    __shared__ Red red;
    __shared__ int sh_i[NW];
    __shared__ int sh_s[8];
    __shared__ int sh_task;
    if (threadIdx.x == 0) {
        const unsigned smid = read_smid();
        s_rot = (NW > 1) ? 32 * (atomicAdd(&sm_slot[smid & 1023], 1) % NW) : 0;
    }
    __syncthreads();
    const int tid = VTID;
    const double q = prm.q_value;
    __shared__ Batch batch;
    while (true) {
      // GHMs without a base cluster in either orientation have nothing to do in phase 2 (their outcome
      // fields were zeroed with the plan); the others are processed 32 queue entries at a time
      pop_batch(batch, items, n_items, next_item, istate, 3, 0, false, tid);
      if (!batch.more) break;
      const int nb = batch.n;
      for (int bi = 0; bi < nb; bi++) {
        GhmTask& task = tasks[batch.item[bi] >> 2];
        if (tid == 0) {
            setup_view(g, task, ws, prm, pool_factor, 0, true, false);
            if (stage > 0) {
                g.nk[0] = task.half_nk[0];
                g.nk[1] = task.half_nk[1];
            }
        }
        __syncthreads();
        if (g.status) {  // pool exhausted in an earlier launch: report, the host retries
            if (tid == 0 && stage == 3) {
                task.status = g.status;
                task.n_mult = 0;
            }
            __syncthreads();
            continue;
        }
        if (stage == 0) {  // orientations too small for phase-1 items are staged here
            const int D0 = (int)task.cnt[0];
            for (int o = 0; o < 2; o++) {
                if ((task.half_mask >> o) & 1) continue;
                const uint32_t* src = dots + task.dot_off + (o ? D0 : 0);
                const int n = (int)task.cnt[o];
                for (int i = tid; i < n; i += NT) {
                    g.key[o][i] = src[i];
                    g.alive[o][i] = 1;
                }
            }
            __syncthreads();
        }
        if (stage == 0 || stage == 2) {
            for (int e = 0; e < 4 && !g.status; e++) {
                const int oc = (e == 1 || e == 2) ? 1 : 0;
                const int om = (e == 1 || e == 3) ? 1 : 0;
                if (g.nbc[oc] > 0 && g.nk[om] > 0) K::enrich_pass(g, prm.gap_size, oc, om, q, sh_i, sh_s);
                __syncthreads();
            }
        } else if (stage == 1) {
            K::filter_pass(g, prm, sh_s);
            __syncthreads();
        } else {
            K::cluster_clusters(g, prm, &red, sh_i, sh_s);
            __syncthreads();
            if (tid == 0) {
                task.status = g.status;
                task.n_mult = g.status ? 0 : sh_s[0];
                task.n_clusters = (g.ncl[0] - g.cl_lo[0]) + (g.ncl[1] - g.cl_lo[1]);
            }
            __syncthreads();
            if (!g.status && task.n_mult > 0) {
                if (hl) pk::measure_ghm<true>(pg, task, ws, ds, pf, pool_factor);
                else pk::measure_ghm<false>(pg, task, ws, ds, pf, pool_factor);
            }
        }
        __syncthreads();
        if (tid == 0 && stage < 3) {
            // every orientation's state now lives in the task record
            task.half_mask = 3;
            for (int o = 0; o < 2; o++) {
                task.half_nbc[o] = g.nbc[o];
                task.half_ncl[o] = g.ncl[o];
                task.half_pool[o] = g.pool_used[o];
                task.half_nk[o] = g.nk[o];
            }
            task.half_status[0] = task.half_status[1] = g.status;
        }
        __syncthreads();
      }
      __syncthreads();
    }
}

}  // namespace

#ifndef IADH_HOST_EMU   // (the host emulation of tests/emu calls the kernels directly)
void launch_cluster(GhmTask* tasks, GhmProf* profs, const int* order, int n_tasks, int* next_task, const uint32_t* dots, uint8_t* ws,
                    const DevParams& prm, int pool_factor, int* sm_slot, const DatasetDev& ds, const ProfileDev& pf, bool hl,
                    int grid, int nt, cudaStream_t st)
{
    if (n_tasks == 0) return;
    switch (nt) {
    case 32: cluster_kernel<32><<<grid, 32, 0, st>>>(tasks, profs, order, n_tasks, next_task, dots, ws, prm, pool_factor, sm_slot, ds, pf, hl ? 1 : 0); break;
    case 64: cluster_kernel<64><<<grid, 64, 0, st>>>(tasks, profs, order, n_tasks, next_task, dots, ws, prm, pool_factor, sm_slot, ds, pf, hl ? 1 : 0); break;
    case 128: cluster_kernel<128><<<grid, 128, 0, st>>>(tasks, profs, order, n_tasks, next_task, dots, ws, prm, pool_factor, sm_slot, ds, pf, hl ? 1 : 0); break;
    default: cluster_kernel<256><<<grid, 256, 0, st>>>(tasks, profs, order, n_tasks, next_task, dots, ws, prm, pool_factor, sm_slot, ds, pf, hl ? 1 : 0); break;
    }
}

// resident kernel: `cfg` = threads per CTA, smem_bytes = dynamic shared memory per CTA (the size class)
static int g_res_attr_done[8] = {0, 0, 0, 0, 0, 0, 0, 0};
template <int NT, int MINB>
static cudaError_t launch_resident_t(int slot, GhmTask* tasks, const int* order, int n_tasks, int* next_task, const uint32_t* dots,
                                     uint8_t* ws, const DevParams& prm, int pool_factor, int smem_bytes, const DatasetDev& ds,
                                     const ProfileDev& pf, bool hl, int grid, cudaStream_t st, unsigned long long* dbg)
{
    if (!g_res_attr_done[slot]) {
        cudaError_t e = cudaFuncSetAttribute(resident_kernel<NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
        if (e != cudaSuccess) return e;
        g_res_attr_done[slot] = 1;
    }
    resident_kernel<NT, MINB><<<grid, NT, smem_bytes, st>>>(tasks, order, n_tasks, next_task, dots, ws, prm, pool_factor, smem_bytes, ds, pf, hl ? 1 : 0, dbg);
    return cudaGetLastError();
}
// resident CTAs of this shape per SM (occupancy of the kernel with `smem_bytes` of dynamic shared memory)
int resident_ctas_per_sm(int nt, int smem_bytes)
{
    int n = 0;
    cudaError_t e;
    switch (nt) {
    case 32:
        cudaFuncSetAttribute(resident_kernel<32, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, resident_kernel<32, 8>, 32, (size_t)smem_bytes);
        break;
    case 64:
        cudaFuncSetAttribute(resident_kernel<64, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, resident_kernel<64, 4>, 64, (size_t)smem_bytes);
        break;
    case 128:
        cudaFuncSetAttribute(resident_kernel<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, resident_kernel<128, 2>, 128, (size_t)smem_bytes);
        break;
    default:
        cudaFuncSetAttribute(resident_kernel<256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, resident_kernel<256, 1>, 256, (size_t)smem_bytes);
        break;
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return 1;
    }
    return n;
}

cudaError_t launch_resident(GhmTask* tasks, const int* order, int n_tasks, int* next_task, const uint32_t* dots, uint8_t* ws,
                            const DevParams& prm, int pool_factor, int smem_bytes, const DatasetDev& ds, const ProfileDev& pf, bool hl,
                            int grid, int nt, cudaStream_t st, unsigned long long* dbg)
{
    if (n_tasks == 0) return cudaSuccess;
    switch (nt) {
    case 32: return launch_resident_t<32, 8>(0, tasks, order, n_tasks, next_task, dots, ws, prm, pool_factor, smem_bytes, ds, pf, hl, grid, st, dbg);
    case 64: return launch_resident_t<64, 4>(1, tasks, order, n_tasks, next_task, dots, ws, prm, pool_factor, smem_bytes, ds, pf, hl, grid, st, dbg);
    case 128: return launch_resident_t<128, 2>(2, tasks, order, n_tasks, next_task, dots, ws, prm, pool_factor, smem_bytes, ds, pf, hl, grid, st, dbg);
    default: return launch_resident_t<256, 1>(3, tasks, order, n_tasks, next_task, dots, ws, prm, pool_factor, smem_bytes, ds, pf, hl, grid, st, dbg);
    }
}

void launch_phase2_stage(GhmTask* tasks, const int* items, int n_items, int* next_item, const uint32_t* dots, uint8_t* ws,
                         const DevParams& prm, int pool_factor, int* sm_slot, const DatasetDev& ds, const ProfileDev& pf,
                         bool hl, int stage, int grid, int nt, cudaStream_t st, const uint32_t* istate)
{
    if (n_items == 0) return;
    #define P2_LAUNCH(NT_, MB_) phase2_stage_kernel<NT_, MB_><<<grid, NT_, 0, st>>>(tasks, items, n_items, next_item, dots, ws, prm, pool_factor, sm_slot, ds, pf, hl ? 1 : 0, stage, istate)
    switch (nt) {
    case 32: P2_LAUNCH(32, 16); break;
    case 3224: P2_LAUNCH(32, 24); break;
    case 128: P2_LAUNCH(128, 4); break;
    case 6412: P2_LAUNCH(64, 12); break;
    case 6416: P2_LAUNCH(64, 16); break;
    case 64: P2_LAUNCH(64, 8); break;
    default: P2_LAUNCH(32, 32); break;   // 3232 (pass_cfg_supported: nothing else gets here)
    }
}

// the (threads, minimum CTAs per SM) pairs the pass-per-launch kernels are instantiated for, as IADH_PASS_CFG
// codes; one table for phase 1 and phase 2 (the engine rejects anything else when the context is created)
bool pass_cfg_supported(int code)
{
    switch (code) {
    case 32: case 3224: case 3232: case 128: case 6412: case 6416: case 64: return true;
    default: return false;
    }
}

// The join pass of the pass-per-launch path runs phase1_join2_kernel (two-phase scan); IADH_JOIN2=0 selects the
// row-by-row join of phase1_pass_kernel instead.  Round 2, on a B200: the whole GPU suite passes with either
// (pass path forced), config 4's join passes take 22.8 ms instead of 49.6 ms (profiles/README.md).
// Read at every launch so a test can switch it per call.
static bool join2_enabled()
{
    const char* e = getenv("IADH_JOIN2");
    return e == nullptr || atoi(e) != 0;
}

void launch_phase1_pass(GhmTask* tasks, const int* items, int n_items, int* next_item, const uint32_t* dots, uint8_t* ws,
                        const DevParams& prm, int pool_factor, int* sm_slot, int pass, int gi, int grid, int nt,
                        cudaStream_t st, unsigned long long* dbg, uint32_t* istate)
{
    if (n_items == 0) return;
    if (pass == 2 && (nt == 32 || nt == 3224 || nt == 3232) && join2_enabled()) {  // IADH_JOIN2=1 (experimental)
        if (nt == 32) phase1_join2_kernel<16><<<grid, 32, 0, st>>>(tasks, items, n_items, next_item, ws, prm, pool_factor, gi, dbg, istate);
        else if (nt == 3224) phase1_join2_kernel<24><<<grid, 32, 0, st>>>(tasks, items, n_items, next_item, ws, prm, pool_factor, gi, dbg, istate);
        else phase1_join2_kernel<32><<<grid, 32, 0, st>>>(tasks, items, n_items, next_item, ws, prm, pool_factor, gi, dbg, istate);
        return;
    }
    // (threads, minimum CTAs per SM): the second number caps the registers per thread
    #define PASS_LAUNCH(NT_, MB_) phase1_pass_kernel<NT_, MB_><<<grid, NT_, 0, st>>>(tasks, items, n_items, next_item, dots, ws, prm, pool_factor, sm_slot, pass, gi, dbg, istate)
    switch (nt) {
    case 32: PASS_LAUNCH(32, 16); break;
    case 3224: PASS_LAUNCH(32, 24); break;
    case 3232: PASS_LAUNCH(32, 32); break;
    case 128: PASS_LAUNCH(128, 4); break;
    case 6412: PASS_LAUNCH(64, 12); break;
    case 6416: PASS_LAUNCH(64, 16); break;
    case 64: PASS_LAUNCH(64, 8); break;
    default: PASS_LAUNCH(32, 32); break;   // 3232
    }
}
#endif  // IADH_HOST_EMU
