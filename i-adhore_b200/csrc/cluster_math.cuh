// Device arithmetic of base clusters: the FP expression trees of BaseCluster.cpp /
// Cluster.cpp restated for sm_100a.  Compiled with --fmad=false: the reference binary is
// plain x86-64 -O3 (no FMA), and every expression below keeps its evaluation order, so
// regression lines, intervals and distances are bit-identical.
#pragma once
#include <cuda_runtime.h>

#include "common.h"
#include "glibc_pow.cuh"

namespace cm {

// Cluster.cpp:16-25
__device__ __forceinline__ int dpd_i(int x1, int y1, int x2, int y2)
{
    int dx = abs(x1 - x2), dy = abs(y1 - y2);
    return dx < dy ? 2 * dy - dx : 2 * dx - dy;
}
// Cluster.cpp:5-14
__device__ __forceinline__ double dpd_d(double x1, double y1, double x2, double y2)
{
    double dx = fabs(x1 - x2), dy = fabs(y1 - y2);
    return dx < dy ? (2 * dy - dx) : (2 * dx - dy);
}

// BaseCluster.cpp:171-179 / 211-220 on exact integer sums (all < 2^53)
__device__ __forceinline__ double r2_sums(long long sx, long long sy, long long sxy, long long sxx,
                                          long long syy, unsigned int n)
{
    const double sum_x = (double)sx, sum_y = (double)sy, sum_xy = (double)sxy, sum_x2 = (double)sxx,
                 sum_y2 = (double)syy;
    const double value = ((sum_x2 - sum_x * sum_x / n) * (sum_y2 - sum_y * sum_y / n));
    if (value == 0) return -1;
    const double r = (sum_xy - (sum_x * sum_y / n)) / sqrt(value);
    return r * r;
}

// BaseCluster.cpp:146-180 with the extra point (X,Y); (0,0) is ignored
__device__ __forceinline__ double r2_with_point(const ClusterRec& c, int X, int Y)
{
    if (X != 0 || Y != 0)
        return r2_sums(c.sx + X, c.sy + Y, c.sxy + (long long)X * Y, c.sxx + (long long)X * X,
                       c.syy + (long long)Y * Y, (unsigned)c.n + 1u);
    return r2_sums(c.sx, c.sy, c.sxy, c.sxx, c.syy, (unsigned)c.n);
}

// BaseCluster.cpp:182-221
__device__ __forceinline__ double r2_merged(const ClusterRec& a, const ClusterRec& b)
{
    return r2_sums(a.sx + b.sx, a.sy + b.sy, a.sxy + b.sxy, a.sxx + b.sxx, a.syy + b.syy,
                   (unsigned)(a.n + b.n));
}

// recompute the integer sums from the AP array
__device__ inline void recompute_sums(ClusterRec& c, const uint32_t* pool)
{
    long long sx = 0, sy = 0, sxy = 0, sxx = 0, syy = 0;
    for (int i = 0; i < c.n; i++) {
        const uint32_t k = pool[c.off + i];
        const long long x = KEY_X(k), y = KEY_Y(k);
        sx += x; sy += y; sxy += x * y; sxx += x * x; syy += y * y;
    }
    c.sx = sx; c.sy = sy; c.sxy = sxy; c.sxx = sxx; c.syy = syy;
}

// BaseCluster::regression (BaseCluster.cpp:266-291) from the exact integer sums
__device__ inline void regress(ClusterRec& c)
{
    const double N = (double)c.n;
    if (N >= 3.0) {
        const double sum_x = (double)c.sx, sum_y = (double)c.sy, sum_xy = (double)c.sxy, sum_x2 = (double)c.sxx;
        c.avg_x = sum_x / N;
        const double avg_y = sum_y / N;
        c.var_x = (sum_x2 - N * c.avg_x * c.avg_x) / (N - 1.0);
        const double cov_xy = (sum_xy - N * c.avg_x * avg_y) / (N - 1.0);
        c.b = cov_xy / c.var_x;
        c.a = avg_y - c.b * c.avg_x;
    }
}

// tail of BaseCluster::updateStatistics (BaseCluster.cpp:258-263) given the two ordered sums
__device__ inline void finish_stats(ClusterRec& c, double mrss, double var_x, const uint32_t* pool)
{
    const int Ni = c.n;
    c.mrss = mrss / (Ni - 2);
    c.var_x = var_x / (Ni - 1);
    c.x_end1 = (double)(unsigned)KEY_X(pool[c.off]);
    c.x_end2 = (double)(unsigned)KEY_X(pool[c.off + Ni - 1]);
    c.y_end1 = c.x_end1 * c.b + c.a;
    c.y_end2 = c.x_end2 * c.b + c.a;
}

// BaseCluster.cpp:238-291 (regression + updateStatistics), executed eagerly after every
// modification (the reference's lazy `b == 0.0` flag yields the same values).
// mrss / var_x are accumulated sequentially in multiset order.
__device__ inline void update_stats(ClusterRec& c, const uint32_t* pool)
{
    regress(c);
    const int Ni = c.n;
    double var_x = 0.0, mrss = 0.0;
    const double a = c.a, b = c.b, avg_x = c.avg_x;
    for (int i = 0; i < Ni; i++) {
        const uint32_t k = pool[c.off + i];
        const int x = KEY_X(k), y = KEY_Y(k);
        const double y_est = b * x + a;
        mrss += (y - y_est) * (y - y_est);
        var_x += (x - avg_x) * (x - avg_x);
    }
    finish_stats(c, mrss, var_x, pool);
}

// BaseCluster.cpp:293-304
__device__ __forceinline__ double dist_to_point(const ClusterRec& c, int x, int y)
{
    if (x >= c.x_end1 && x <= c.x_end2) return 0;
    if ((double)x < c.x_end1) return dpd_d((double)x, (double)y, c.x_end1, c.y_end1);
    return dpd_d((double)x, (double)y, c.x_end2, c.y_end2);
}

// BaseCluster.cpp:306-336
__device__ __forceinline__ bool in_interval(const ClusterRec& c, int x, int y)
{
    const double n = (double)c.n;
    const double pi = 1.96 * sqrt(c.mrss) * sqrt(1.0 + 1.0 / n + (x - c.avg_x) * (x - c.avg_x) / (n - 1) / c.var_x);
    const double up = (c.a + c.b * x) + pi;
    const double down = (c.a + c.b * x) - pi;
    return y >= down && y <= up;
}

// BaseCluster.cpp:353-364
__device__ __forceinline__ bool intersect(double x1, double y1, double x2, double y2, double x3, double y3,
                                          double x4, double y4)
{
    const double ud = (y4 - y3) * (x2 - x1) - (x4 - x3) * (y2 - y1);
    if (ud != 0) {
        const double ua = ((x4 - x3) * (y1 - y3) - (y4 - y3) * (x1 - x3)) / ud;
        const double ub = ((x2 - x1) * (y1 - y3) - (y2 - y1) * (x1 - x3)) / ud;
        if (ua >= 0 && ua <= 1 && ub >= 0 && ub <= 1) return true;
    }
    return false;
}
__device__ __forceinline__ double mind(double a, double b) { return (a < b) ? a : b; }
__device__ __forceinline__ double maxd(double a, double b) { return (a > b) ? a : b; }

// BaseCluster.cpp:376-396
__device__ __forceinline__ double point_line_dpd(double x1, double y1, double x2, double y2, double x3, double y3)
{
    const double dx = x2 - x1, dy = y2 - y1;
    double x0, y0;
    if ((dx == 0) && (dy == 0)) {
        x0 = x1;
        y0 = y1;
    } else {
        double t = ((x3 - x1) * dx + (y3 - y1) * dy) / (dx * dx + dy * dy);
        t = mind(maxd(0, t), 1);
        x0 = x1 + t * dx;
        y0 = y1 + t * dy;
    }
    return dpd_d(x0, y0, x3, y3);
}

// BaseCluster.cpp:398-465.  s*/e* = first / last anchor point (multiset order) of each cluster.
__device__ inline double dist_to_cluster(const ClusterRec& A, const ClusterRec& B, uint32_t s1k, uint32_t e1k,
                                         uint32_t s2k, uint32_t e2k)
{
    if (intersect(A.x_end1, A.y_end1, A.x_end2, A.y_end2, B.x_end1, B.y_end1, B.x_end2, B.y_end2)) return 0;
    double lo = point_line_dpd(A.x_end1, A.y_end1, A.x_end2, A.y_end2, B.x_end1, B.y_end1);
    lo = mind(lo, point_line_dpd(A.x_end1, A.y_end1, A.x_end2, A.y_end2, B.x_end2, B.y_end2));
    lo = mind(lo, point_line_dpd(B.x_end1, B.y_end1, B.x_end2, B.y_end2, A.x_end1, A.y_end1));
    lo = mind(lo, point_line_dpd(B.x_end1, B.y_end1, B.x_end2, B.y_end2, A.x_end2, A.y_end2));
    const int s1x = KEY_X(s1k), s1y = KEY_Y(s1k), e1x = KEY_X(e1k), e1y = KEY_Y(e1k);
    const int s2x = KEY_X(s2k), s2y = KEY_Y(s2k), e2x = KEY_X(e2k), e2y = KEY_Y(e2k);
    lo = mind(lo, (double)dpd_i(s1x, s1y, s2x, s2y));
    lo = mind(lo, (double)dpd_i(s1x, s1y, e2x, e2y));
    lo = mind(lo, (double)dpd_i(e1x, e1y, s2x, s2y));
    lo = mind(lo, (double)dpd_i(e1x, e1y, e2x, e2y));
    lo = mind(lo, point_line_dpd(A.x_end1, A.y_end1, A.x_end2, A.y_end2, s2x, s2y));
    lo = mind(lo, point_line_dpd(A.x_end1, A.y_end1, A.x_end2, A.y_end2, e2x, e2y));
    lo = mind(lo, point_line_dpd(B.x_end1, B.y_end1, B.x_end2, B.y_end2, s1x, s1y));
    lo = mind(lo, point_line_dpd(B.x_end1, B.y_end1, B.x_end2, B.y_end2, e1x, e1y));
    return lo;
}

// first index in [0,n) whose x is >= x / > x (the AP array is x-sorted)
__device__ __forceinline__ int ap_lower(const uint32_t* ap, int n, int x)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (KEY_X(ap[mid]) < x) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ int ap_upper(const uint32_t* ap, int n, int x)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (KEY_X(ap[mid]) <= x) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// BaseCluster.cpp:491-546: every AP of B inside A's x-range must lie in A's prediction interval and
// vice versa.  The reference scans all APs and tests the x-range per AP; the arrays are x-sorted, so
// the same APs are reached by two binary searches (same tests, same order, same result).
__device__ inline bool partial_overlap(const ClusterRec& A, const ClusterRec& B, const uint32_t* pool)
{
    const uint32_t* pa = pool + A.off;
    const uint32_t* pb = pool + B.off;
    const int x1min = KEY_X(pa[0]), x1max = KEY_X(pa[A.n - 1]);
    const int x2min = KEY_X(pb[0]), x2max = KEY_X(pb[B.n - 1]);
    if (x2max >= x1min && x2min <= x1max) {
        const int lo = ap_lower(pb, B.n, x1min), hi = ap_upper(pb, B.n, x1max);
        for (int i = lo; i < hi; i++)
            if (!in_interval(A, KEY_X(pb[i]), KEY_Y(pb[i]))) return false;
        const int lo2 = ap_lower(pa, A.n, x2min), hi2 = ap_upper(pa, A.n, x2max);
        for (int i = lo2; i < hi2; i++)
            if (!in_interval(B, KEY_X(pa[i]), KEY_Y(pa[i]))) return false;
    }
    return true;
}

// ---------------------------------------------------------------------------------------
// Fallback power for bases outside glibc_pow_pos()'s domain (only reachable through the omxn
// hook): correctly rounded double-double square-and-multiply.
struct dd {
    double hi, lo;
};
__device__ __forceinline__ dd two_prod(double a, double b)
{
    dd r;
    r.hi = a * b;
    r.lo = __fma_rn(a, b, -r.hi);
    return r;
}
__device__ __forceinline__ dd dd_mul(dd a, dd b)
{
    dd p = two_prod(a.hi, b.hi);
    p.lo += a.hi * b.lo + a.lo * b.hi;
    dd r;
    r.hi = p.hi + p.lo;
    r.lo = p.lo - (r.hi - p.hi);
    return r;
}
__device__ inline double pow_int(double base, int n)
{
    if (n == 0) return 1.0;
    dd r = {1.0, 0.0}, b = {base, 0.0};
    unsigned e = (unsigned)n;
    while (true) {
        if (e & 1u) r = dd_mul(r, b);
        e >>= 1;
        if (!e) break;
        b = dd_mul(b, b);
    }
    return r.hi + r.lo;
}

// pow(base, n) as the reference's libm computes it (BaseCluster.cpp:567, hpmath.cpp:21)
__device__ inline double pow_ref(double base, int n)
{
    if (n == 0) return 1.0;
    if (base > 0.0 && base < 1e300 && n > 0 && base >= 2.2250738585072014e-308) return glibc_pow_pos(base, (double)n);
    if (base == 0.0 && n > 0) return 0.0;
    return pow_int(base, n);
}

// hpmath.cpp:18-36
__device__ inline double omxn(double x, int n)
{
    if (fabs(x) > 0.1) return pow_ref(1.0 - x, n);
    const double epsilon = fabs(x * 2.2204460492503131e-16);
    double term = 1.0, result = term;
    int i = 0;
    while ((fabs(term) > epsilon) && i != n) {
        i++;
        term = -term * (n - i + 1.0) / (i * 1.0) * x;
        result += term;
    }
    return result;
}

}  // namespace cm
