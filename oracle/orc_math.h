/* oracle/orc_math.h -- the elementary functions of the CPU oracle, written out as explicit sequences of
 * IEEE-754 operations (+ - * / sqrt fma rint), so that a second implementation of the same sequences
 * (the CUDA kernels) can be compared BIT FOR BIT instead of "to a few ulp".
 *
 * TEST INFRASTRUCTURE ONLY (see toss_oracle.h).  libm's log/atan/atan2/sin/cos/pow are correct to < 1 ulp
 * but not reproducible across libraries; these are plain Taylor series on reduced arguments, accurate to
 * ~2 ulp, and are themselves pinned against libm in tests/test_oracle_math.py.  Build with
 * -ffp-contract=off: every fused multiply-add is spelled ORC_FMA, nothing else may be fused.
 */
#ifndef ORC_MATH_H
#define ORC_MATH_H
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "orc_math_constants.h"

#define ORC_FMA(a, b, c) __builtin_fma((a), (b), (c))

/* 2 atanh(z) = 2z (1 + z^2 (1/3 + z^2/5 + ...)), nterms coefficients: 11 for z*z < 0.04, 5 for z*z < 0.0025 */
static inline double orc_two_atanh_series_n(double z, double z2, int nterms)
{
    double p = ORC_ATANH[nterms - 1];
    for (int k = nterms - 2; k >= 0; --k) p = ORC_FMA(p, z2, ORC_ATANH[k]);
    const double zz = z + z;
    return ORC_FMA(zz * z2, p, zz);
}
/* the common path of the face-pair sum: minimax polynomials (weight z^2), 7 coefficients for z*z <= 0.0401 and
 * 5 for t*t <= 0.01005 (tools/gen_math_constants.py); approximation error 1.6e-17 / 4.2e-17 of the value */
static inline double orc_two_atanh_mm(double z, double z2)
{
    double p = ORC_ATANH_MM[6];
    for (int k = 5; k >= 0; --k) p = ORC_FMA(p, z2, ORC_ATANH_MM[k]);
    const double zz = z + z;
    return ORC_FMA(zz * z2, p, zz);
}
static inline double orc_two_atan_mm(double t, double t2) /* 2 atan(t) */
{
    double p = ORC_ATAN_MM[4];
    for (int k = 3; k >= 0; --k) p = ORC_FMA(p, t2, ORC_ATAN_MM[k]);
    const double tt = t + t;
    return ORC_FMA(tt * t2, p, tt);
}
/* far tier of the face-pair sum: 4 coefficients for z*z, t*t <= 0.0025 (approximation error 2.0e-17 of the value) */
static inline double orc_two_atanh_far(double z, double z2)
{
    double p = ORC_ATANH_FAR[3];
    for (int k = 2; k >= 0; --k) p = ORC_FMA(p, z2, ORC_ATANH_FAR[k]);
    const double zz = z + z;
    return ORC_FMA(zz * z2, p, zz);
}
/* the remaining rows of the lane-per-pair kernels: 11 coefficients for z*z <= 0.1601 (error 4.0e-18) */
static inline double orc_two_atanh_wide(double z, double z2)
{
    double p = ORC_ATANH_WIDE[10];
    for (int k = 9; k >= 0; --k) p = ORC_FMA(p, z2, ORC_ATANH_WIDE[k]);
    const double zz = z + z;
    return ORC_FMA(zz * z2, p, zz);
}
/* middle tier: 5 coefficients for z*z <= 0.01005 (approximation error 4.6e-17 of the value) */
static inline double orc_two_atanh_mid(double z, double z2)
{
    double p = ORC_ATANH_MID[4];
    for (int k = 3; k >= 0; --k) p = ORC_FMA(p, z2, ORC_ATANH_MID[k]);
    const double zz = z + z;
    return ORC_FMA(zz * z2, p, zz);
}
static inline double orc_two_atan_far(double t, double t2)
{
    double p = ORC_ATAN_FAR[3];
    for (int k = 2; k >= 0; --k) p = ORC_FMA(p, t2, ORC_ATAN_FAR[k]);
    const double tt = t + t;
    return ORC_FMA(tt * t2, p, tt);
}
static inline double orc_two_atanh_series(double z) { return orc_two_atanh_series_n(z, z * z, 11); }

/* ln(x), x > 0 normal: x = 2^e m, m in [sqrt(1/2), sqrt(2)); ln m = 2 atanh((m-1)/(m+1)) */
static inline double orc_log(double x)
{
    uint64_t b;
    memcpy(&b, &x, 8);
    int e = (int)((b >> 52) & 0x7ff) - 1023;
    b = (b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
    double m;
    memcpy(&m, &b, 8);
    if (m > ORC_SQRT2) {
        m = m * 0.5;
        e += 1;
    }
    const double s = (m - 1.0) / (m + 1.0);
    const double L = orc_two_atanh_series(s);
    const double de = (double)e;
    return ORC_FMA(de, ORC_LN2_HI, ORC_FMA(de, ORC_LN2_LO, L));
}

/* 2 atanh(z), 0 <= z < 1 */
static inline double orc_two_atanh(double z)
{
    if (z * z < 0.04) return orc_two_atanh_series(z);
    return orc_log((1.0 + z) / (1.0 - z));
}

/* atan(t) = t (1 + t^2 (-1/3 + t^2/5 - ...)), nterms coefficients: 22 for |t| <= tan(pi/8), 8 for |t| <= 0.1,
 * 3 for t*t < 1e-4 */
static inline double orc_atan_series_n(double t, double t2, int nterms)
{
    double p = ORC_ATAN[nterms - 1];
    for (int k = nterms - 2; k >= 0; --k) p = ORC_FMA(p, t2, ORC_ATAN[k]);
    return ORC_FMA(t * t2, p, t);
}
static inline double orc_atan_series(double t, int nterms) { return orc_atan_series_n(t, t * t, nterms); }

/* atan2(y, x): a = min/max in [0,1]; a > tan(pi/8) -> (a-1)/(a+1) + pi/4; 22-term series on |t| <= tan(pi/8) */
static inline double orc_atan2(double y, double x)
{
    const double ax = fabs(x), ay = fabs(y);
    const double mx = ay > ax ? ay : ax, mn = ay > ax ? ax : ay;
    double r = 0.0;
    if (mx != 0.0) {
        const double a = mn / mx;
        double t = a, base = 0.0;
        if (a > ORC_TAN_PIO8) {
            t = (a - 1.0) / (a + 1.0);
            base = ORC_PIO4;
        }
        r = base + orc_atan_series(t, 22);
    }
    if (ay > ax) r = ORC_PIO2 - r;
    if (x < 0.0) r = ORC_PI - r;
    return y < 0.0 ? -r : r;
}

/* sin and cos of th, |th| < ~1e5: Cody-Waite reduction by pi/2 with fma, 9-term series on |r| <= pi/4 */
static inline void orc_sincos(double th, double *s, double *c)
{
    const double n = rint(th * ORC_TWO_OVER_PI);
    double r = ORC_FMA(-n, ORC_PIO2_HI, th);
    r = ORC_FMA(-n, ORC_PIO2_LO, r);
    const double r2 = r * r;
    double sp = ORC_SIN[8], cp = ORC_COS[8];
    for (int k = 7; k >= 0; --k) {
        sp = ORC_FMA(sp, r2, ORC_SIN[k]);
        cp = ORC_FMA(cp, r2, ORC_COS[k]);
    }
    const double ss = ORC_FMA(r * r2, sp, r);
    const double cc = ORC_FMA(r2, cp, 1.0);
    switch ((long long)n & 3) {
    case 0: *s = ss; *c = cc; break;
    case 1: *s = cc; *c = -ss; break;
    case 2: *s = -ss; *c = -cc; break;
    default: *s = -cc; *c = ss; break;
    }
}

/* x^(1/8) and x^(1/4) by correctly rounded square roots (desolver's (tol/err)^(1/8); the fitness' ^0.25) */
static inline double orc_root8(double x) { return sqrt(sqrt(sqrt(x))); }
static inline double orc_root4(double x) { return sqrt(sqrt(x)); }

/* sum of 32 lane values in the order of an xor-butterfly (16, 8, 4, 2, 1): what __shfl_xor reductions compute */
static inline double orc_butterfly32(double *v)
{
    for (int o = 16; o > 0; o >>= 1) {
        double t[32];
        for (int i = 0; i < 32; ++i) t[i] = v[i] + v[i ^ o];
        memcpy(v, t, sizeof(t));
    }
    return v[0];
}

/* sum of n values in the order of a 256-thread block: thread t adds items t, t+256, ...; each warp of 32
 * threads butterfly-reduces; the 8 warp sums are added left to right. */
typedef double (*orc_item_fn)(int64_t i, void *arg);
static inline double orc_block_sum256(int64_t n, orc_item_fn item, void *arg)
{
    double part[256];
    for (int t = 0; t < 256; ++t) {
        double s = 0.0;
        for (int64_t i = t; i < n; i += 256) s += item(i, arg);
        part[t] = s;
    }
    double r = 0.0;
    for (int w = 0; w < 8; ++w) r += orc_butterfly32(part + 32 * w);
    return r;
}
#endif
