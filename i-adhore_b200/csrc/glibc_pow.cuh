// pow(x, y) exactly as glibc 2.39's x86-64 FMA variant (__pow_fma, selected by the libm ifunc
// on every AVX2+FMA host) computes it, for the domain the hot path needs:
//   x normal, positive, finite;  y = a positive integer < 2^63 (BaseCluster.cpp:562-567 passes
//   an int); results that are not tiny (|y*log x| < 512) -- below that 1 - pow() is 1.0 anyway.
// Restated operation by operation from the algorithm of sysdeps/ieee754/dbl-64/e_pow.c
// (log_inline / exp_inline, tables in glibc_pow_tables.h): every fused multiply-add the
// compiled routine executes is an explicit fma() here and every other product and sum is a
// separately rounded operation (--fmad=false on the device), so the result is bit-identical
// to the reference's libm call.  tests/test_gpu_parity.py::test_pow_matches_host_libm and
// tests/test_pow_port.py (host compile of this same header) pin it.
#pragma once
#include <stdint.h>
#include <string.h>

#include "glibc_pow_tables.h"

#ifdef __CUDACC__
#define GP_FN __device__ __forceinline__
#define GP_FMA(a, b, c) __fma_rn((a), (b), (c))
#define GP_ASU64(d) ((uint64_t)__double_as_longlong(d))
#define GP_ASDBL(u) __longlong_as_double((long long)(u))
#else
#include <math.h>
#define GP_FN static inline
#define GP_FMA(a, b, c) fma((a), (b), (c))
static inline uint64_t GP_ASU64(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static inline double GP_ASDBL(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
#endif

GP_FN double glibc_pow_pos(double x, double y)
{
    // ---- log_inline: (hi, lo) ~ log(x) with ~68 bits
    const uint64_t ix = GP_ASU64(x);
    const uint64_t tmp = ix - 0x3fe6955500000000ull;
    const int i = (int)((tmp >> 45) & 0x7f);
    const int k = (int)((int64_t)tmp >> 52);
    const uint64_t iz = ix - (tmp & 0xfff0000000000000ull);
    const double z = GP_ASDBL(iz);
    const double kd = (double)k;
    const double invc = gp_log_tab[3 * i], logc = gp_log_tab[3 * i + 1], logctail = gp_log_tab[3 * i + 2];
    const double r = GP_FMA(z, invc, -1.0);
    const double t1 = GP_FMA(kd, GP_LN2HI, logc);
    const double t2 = t1 + r;
    const double lo1 = GP_FMA(kd, GP_LN2LO, logctail);
    const double lo2 = (t1 - t2) + r;
    const double ar = GP_A0 * r;
    const double ar2 = r * ar;
    const double ar3 = r * ar2;
    const double hi = t2 + ar2;
    const double lo3 = GP_FMA(ar, r, -ar2);
    const double lo4 = (t2 - hi) + ar2;
    const double q12 = GP_FMA(r, GP_A2, GP_A1);
    const double q34 = GP_FMA(r, GP_A4, GP_A3);
    const double q56 = GP_FMA(r, GP_A6, GP_A5);
    const double qa = GP_FMA(q56, ar2, q34);
    const double q = GP_FMA(ar2, qa, q12);
    double lo = lo1 + lo2;
    lo = lo + lo3;
    lo = lo + lo4;
    lo = GP_FMA(ar3, q, lo);
    const double lhi = hi + lo;
    const double llo = (hi - lhi) + lo;
    // ---- y * log(x) as ehi + elo
    const double ehi = y * lhi;
    const double elo = GP_FMA(y, llo, GP_FMA(lhi, y, -ehi));
    // ---- exp_inline(ehi, elo)
    const uint32_t abstop = (uint32_t)((GP_ASU64(ehi) >> 52) & 0x7ff);
    if (abstop - 0x3c9u > 0x3eu) {
        if ((int32_t)(abstop - 0x3c9u) < 0) return 1.0 + ehi;  // |ehi| < 2^-54
        // |ehi| >= 512: x^y < e^-512 (or overflow, impossible for x <= 1): far below 2^-54, the
        // caller's 1 - pow() is exactly 1.0 whatever is returned here
        return ehi < 0 ? 0.0 : GP_ASDBL(0x7ff0000000000000ull);
    }
    const double zz = GP_FMA(ehi, GP_INVLN2N, GP_SHIFT);
    const uint64_t ki = GP_ASU64(zz);
    const double kd2 = zz - GP_SHIFT;
    double rr = GP_FMA(kd2, GP_NEGLN2HIN, ehi);
    rr = GP_FMA(kd2, GP_NEGLN2LON, rr);
    rr = elo + rr;
    const int idx = (int)(ki & 0x7f);
    const uint64_t sbits = gp_exp_tab[2 * idx + 1] + (ki << 45);
    const double tail = GP_ASDBL(gp_exp_tab[2 * idx]);
    const double c23 = GP_FMA(rr, GP_C3, GP_C2);
    const double tr = rr + tail;
    const double r2 = rr * rr;
    const double c45 = GP_FMA(rr, GP_C5, GP_C4);
    double t = GP_FMA(c23, r2, tr);
    const double r4 = r2 * r2;
    t = GP_FMA(c45, r4, t);
    const double scale = GP_ASDBL(sbits);
    return GP_FMA(t, scale, scale);
}
