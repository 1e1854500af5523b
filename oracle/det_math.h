/*
 * oracle/det_math.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Deterministic replacements for the libm transcendentals that libmypaint's
 * brush engine calls (expf/exp/log/hypotf).  They use only IEEE-754 basic
 * operations (+ - * / sqrt fma, int<->double conversions), every one of which
 * is correctly rounded on x86-64 (gcc -ffp-contract=off) and on sm_100a
 * (nvcc --fmad=false, explicit fma()), so the CPU oracle in "det" mode and the
 * CUDA rasteriser produce BIT-IDENTICAL brush-state trajectories.  The CUDA
 * side has its own independently written copy
 * (spiral-tensorflow_b200/csrc/detmath.cuh); tests/ compares the two.
 *
 * Accuracy: both functions are evaluated in double to <= ~4e-15 relative
 * error, i.e. after rounding to float they agree with a correctly rounded
 * expf/logf except when the exact value lies within ~1e-7 ulp of a rounding
 * boundary.  The oracle's "libm" mode (glibc) is the upstream-faithful one;
 * tests/test_oracle_paint.py measures how often the two modes differ.
 *
 * SPEC (shared with the CUDA copy; change both or neither):
 *   det_exp(x):  k = rint(x*LOG2E); r = fma(-k,LN2_HI,x); r = fma(-k,LN2_LO,r);
 *                p = Estrin degree-12 Taylor of exp(r) (fma tree below);
 *                result = p * 2^(k/2) * 2^(k-k/2)     (exact scalings)
 *   det_log(x):  x = m*2^e, m in [1,2); if m > SQRT2 {m*=0.5; e+=1}
 *                t = (m-1)/(m+1); w = t*t;
 *                q = Estrin degree-8 poly in w with c_k = 2/(2k+1)
 *                result = fma(e, LN2_HI, fma(t, q, e*LN2_LO))
 *   det_hypotf(a,b) = (float) sqrt((double)a*a + (double)b*b)
 */
#ifndef SPIRAL_ORACLE_DET_MATH_H
#define SPIRAL_ORACLE_DET_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#define DM_LOG2E   0x1.71547652b82fep+0
#define DM_LN2_HI  0x1.62e42fee00000p-1
#define DM_LN2_LO  0x1.a39ef35793c76p-33
#define DM_SQRT2   0x1.6a09e667f3bcdp+0

static inline double dm_from_bits(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static inline uint64_t dm_to_bits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }

/* 2^k for k in [-1022, 1023] */
static inline double dm_pow2i(int k) { return dm_from_bits((uint64_t)(k + 1023) << 52); }

static inline double det_exp(double x)
{
    if (x != x) return x;
    if (x > 709.0) return INFINITY;
    if (x < -745.0) return 0.0;
    double k = rint(x * DM_LOG2E);
    double r = fma(-k, DM_LN2_HI, x);
    r = fma(-k, DM_LN2_LO, r);
    /* Taylor coefficients 1/j!, j = 0..12 */
    const double c2 = 0x1.0000000000000p-1, c3 = 0x1.5555555555555p-3;
    const double c4 = 0x1.5555555555555p-5, c5 = 0x1.1111111111111p-7;
    const double c6 = 0x1.6c16c16c16c17p-10, c7 = 0x1.a01a01a01a01ap-13;
    const double c8 = 0x1.a01a01a01a01ap-16, c9 = 0x1.71de3a556c734p-19;
    const double c10 = 0x1.27e4fb7789f5cp-22, c11 = 0x1.ae64567f544e4p-26;
    const double c12 = 0x1.1eed8eff8d898p-29;
    double r2 = r * r;
    double r4 = r2 * r2;
    double r8 = r4 * r4;
    double p01 = fma(1.0, r, 1.0);      /* c0 + c1 r */
    double p23 = fma(c3, r, c2);
    double p45 = fma(c5, r, c4);
    double p67 = fma(c7, r, c6);
    double p89 = fma(c9, r, c8);
    double pab = fma(c11, r, c10);
    double q0 = fma(p23, r2, p01);      /* degree 0..3 */
    double q1 = fma(p67, r2, p45);      /* degree 4..7  (/r^4) */
    double q2 = fma(pab, r2, p89);      /* degree 8..11 (/r^8) */
    q2 = fma(c12, r4, q2);              /* + degree 12 */
    double p = fma(q1, r4, q0);
    p = fma(q2, r8, p);
    int ki = (int)k;
    int k1 = ki / 2;          /* C truncation toward zero */
    int k2 = ki - k1;
    return p * dm_pow2i(k1) * dm_pow2i(k2);
}

/* x must be finite, > 0 and a normal double (always true for promoted floats > 0) */
static inline double det_log(double x)
{
    uint64_t b = dm_to_bits(x);
    int e = (int)((b >> 52) & 0x7ff) - 1023;
    double m = dm_from_bits((b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
    if (m > DM_SQRT2) { m = m * 0.5; e += 1; }
    double t = (m - 1.0) / (m + 1.0);
    double w = t * t;
    const double c0 = 0x1.0000000000000p+1, c1 = 0x1.5555555555555p-1;
    const double c2 = 0x1.999999999999ap-2, c3 = 0x1.2492492492492p-2;
    const double c4 = 0x1.c71c71c71c71cp-3, c5 = 0x1.745d1745d1746p-3;
    const double c6 = 0x1.3b13b13b13b14p-3, c7 = 0x1.1111111111111p-3;
    const double c8 = 0x1.e1e1e1e1e1e1ep-4;
    double w2 = w * w;
    double w4 = w2 * w2;
    double w8 = w4 * w4;
    double p01 = fma(c1, w, c0);
    double p23 = fma(c3, w, c2);
    double p45 = fma(c5, w, c4);
    double p67 = fma(c7, w, c6);
    double q0 = fma(p23, w2, p01);
    double q1 = fma(p67, w2, p45);
    double q = fma(q1, w4, q0);
    q = fma(c8, w8, q);
    double ed = (double)e;
    return fma(ed, DM_LN2_HI, fma(t, q, ed * DM_LN2_LO));
}

static inline float det_hypotf(float a, float b)
{
    double da = (double)a, db = (double)b;
    return (float)sqrt(da * da + db * db);
}

#endif
