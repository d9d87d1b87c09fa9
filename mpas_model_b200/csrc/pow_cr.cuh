// pow_cr.cuh -- x**y for the exner function (TI:3332, TI:6879) evaluated to ~2^-68 before the final rounding.
//
// Why: the reference's fp64 CPU build gets x**rcv from glibc's pow(), which is correctly rounded in all but ~1 case in
// 10^3..10^4; CUDA's pow() is a 2-ulp function.  The 1-ulp differences in exner are harmless for every prognostic but w:
// pressure_p = zz*rgas*(exner*rtheta_p + rtheta_base*(exner - exner_base)) cancels 3 digits and the pressure-gradient /
// buoyancy residual that drives w cancels 5 more, so an ulp in exner shows up at 1e-11 of |w| after ONE step -- the
// north-star bar itself.  An (almost always) correctly rounded pow on the device makes exner bit-identical to the CPU
// arithmetic in > 99.9 % of the points instead of ~85 %.
//
// Method: log in double-double from a 128-entry table (x = 2^e * m, z = m * invc_i - 1 exactly, log1p(z) by series),
// y * log(x) in double-double, exp from a 128-entry table of 2^(j/128) and a short series.  Only the cases the dycore
// produces take this path (x, result normal and positive); everything else falls back to pow().
#pragma once
#include <cmath>
#ifdef __CUDACC__
#define POWCR_HD __device__ __forceinline__
#else
#define POWCR_HD inline
#define __device__
#endif
#include "pow_cr_tables.inl"

namespace powcr {
struct dd { double h, l; };
POWCR_HD dd two_sum(double a, double b) { const double s = a + b, bb = s - a; dd r; r.h = s; r.l = (a - (s - bb)) + (b - bb); return r; }
POWCR_HD dd fast_two_sum(double a, double b) { const double s = a + b; dd r; r.h = s; r.l = b - (s - a); return r; }     // |a| >= |b|
POWCR_HD dd two_prod(double a, double b) { dd r; r.h = a * b; r.l = fma(a, b, -r.h); return r; }
POWCR_HD dd add(dd a, dd b) { dd s = two_sum(a.h, b.h); s.l += a.l + b.l; return fast_two_sum(s.h, s.l); }
POWCR_HD dd add(dd a, double b) { dd s = two_sum(a.h, b); s.l += a.l; return fast_two_sum(s.h, s.l); }
}

POWCR_HD double pow_cr(double x, double y) {
    using namespace powcr;
#ifdef __CUDA_ARCH__
    long long ix = __double_as_longlong(x);
#else
    long long ix; memcpy(&ix, &x, 8);
#endif
    const int bexp = (int)((ix >> 52) & 0x7ff);
    if (ix <= 0 || bexp == 0 || bexp == 0x7ff || !(fabs(y) < 1.0e3)) return pow(x, y);       // not a positive normal number
    const int e = bexp - 1023;
    const int i = (int)((ix >> 45) & 127);
    const long long im = (ix & 0x000fffffffffffffLL) | 0x3ff0000000000000LL;
#ifdef __CUDA_ARCH__
    const double m = __longlong_as_double(im);
#else
    double m; memcpy(&m, &im, 8);
#endif
    const double invc = POWCR_LOG[i][0];
    // z = m * invc - 1 exactly, as a double-double
    const dd p = two_prod(m, invc);
    const dd z = fast_two_sum(p.h - 1.0, p.l);
    // log1p(z) = z - z^2/2 + z^3 (1/3 - z/4 + ...), |z| < 2^-7.9
    const dd zz = two_prod(z.h, z.h);
    dd s2; s2.h = -0.5 * zz.h; s2.l = -0.5 * (zz.l + 2.0 * z.h * z.l);
    const double zh = z.h;
    const double t = zh * zh * zh * (1.0 / 3 + zh * (-0.25 + zh * (0.2 + zh * (-1.0 / 6 + zh * (1.0 / 7 + zh * (-0.125 + zh * (1.0 / 9)))))));
    const double ed = (double)e;
    dd L; L.h = ed * POWCR_LN2_H; L.l = ed * POWCR_LN2_M;      // the head product is exact
    dd lc; lc.h = POWCR_LOG[i][1]; lc.l = POWCR_LOG[i][2] + ed * POWCR_LN2_L;
    L = add(L, lc);
    L = add(L, z);
    L = add(L, s2);
    L = add(L, t);
    // E = y * L
    const dd yp = two_prod(y, L.h);
    const dd E = fast_two_sum(yp.h, yp.l + y * L.l);
    if (!(fabs(E.h) < 700.0)) return pow(x, y);
    // exp(E): E = k * ln2/128 + r
    const double kd = rint(E.h * POWCR_INVC);
    const long long k = (long long)kd;
    const double r0 = fma(-kd, POWCR_C_H, E.h);                // exact
    const dd r = two_sum(r0, E.l - kd * POWCR_C_M - kd * POWCR_C_L);
    const double rh = r.h;
    const double q = r.l + rh * rh * (0.5 + rh * (1.0 / 6 + rh * (1.0 / 24 + rh * (1.0 / 120 + rh * (1.0 / 720 + rh * (1.0 / 5040))))));
    const dd w = two_sum(rh, q);                                // exp(r) - 1
    const int j = (int)(k & 127);
    const double th = POWCR_EXP[j][0], tl = POWCR_EXP[j][1];
    const dd pw = two_prod(th, w.h);
    const double small = pw.l + tl + th * w.l + tl * w.h;
    const dd s = two_sum(th, pw.h);
    const double res = s.h + (s.l + small);
    const long long sc = (k >> 7);                              // floor division: the table index is k mod 128 >= 0
#ifdef __CUDA_ARCH__
    long long ir = __double_as_longlong(res);
#else
    long long ir; memcpy(&ir, &res, 8);
#endif
    const long long rexp = ((ir >> 52) & 0x7ff) + sc;
    if (rexp <= 0 || rexp >= 0x7ff) return pow(x, y);           // result not normal
    ir += sc << 52;
#ifdef __CUDA_ARCH__
    return __longlong_as_double(ir);
#else
    double out; memcpy(&out, &ir, 8); return out;
#endif
}
// PRECISION=single: powf of the reference's libm is correctly rounded (it is evaluated in double); so is this
POWCR_HD float pow_cr(float x, float y) { return (float)pow((double)x, (double)y); }
