/*
 * oracle/fast_math.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * binary32 exp / log / hypot of the rasteriser's FAST arithmetic mode (PO_MATH_FAST): fixed expression
 * trees of IEEE-754 single-precision operations (+ - * fmaf sqrtf, bit moves), every one correctly rounded
 * on x86-64 (gcc -ffp-contract=off; fmaf() is exact whether glibc dispatches it to the FMA unit or emulates
 * it) and on sm_100a (nvcc --fmad=false), so the oracle in "fast" mode and the CUDA rasteriser produce
 * BIT-IDENTICAL brush trajectories.  The CUDA side has its own copy
 * (spiral-tensorflow_b200/csrc/fastmath.cuh); tests/ compares the two.
 *
 * Accuracy: within ~2 ulp of glibc's expf / logf / hypotf (tests/test_oracle_paint.py).  The "libm" mode is
 * the upstream-faithful one; the tests measure how far FAST canvases are from it.
 *
 * SPEC (shared with the CUDA copy; change both or neither):
 *   exp_pm(x):  x = clamp(x, -86, 86); t = fmaf(x, LOG2E, 1.5*2^23); k = t - 1.5*2^23;
 *               r = fmaf(-k, LN2_HI, x); r = fmaf(-k, LN2_LO, r); r2 = r*r; r4 = r2*r2;
 *               A = fmaf(1/720, r4, fmaf(1/24, r2, 1/2));  B = fmaf(1/5040, r4, fmaf(1/120, r2, 1/6));
 *               exp(x)  = bits(1 + fmaf(r2, fmaf( r, B, A),  r)) + (k << 23)
 *               exp(-x) = bits(1 + fmaf(r2, fmaf(-r, B, A), -r)) - (k << 23)
 *   logf(x):    i = bits(x) - bits(sqrt(.5)); e = i >> 23; m = x * 2^-e in [sqrt(.5), sqrt(2)); f = m - 1;
 *               log m = f - f^2/2 + f^3 R(f), R of degree 6 (Estrin); result = e*LN2_HI + (e*LN2_LO + log m)
 *   hypotf(a,b) = sqrtf(fmaf(a, a, b*b))
 */
#ifndef SPIRAL_ORACLE_FAST_MATH_H
#define SPIRAL_ORACLE_FAST_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#define FM_LOG2E   0x1.715476p+0f
#define FM_LN2_HI  0x1.62e400p-1f
#define FM_LN2_LO  0x1.7f7d1cp-20f
#define FM_MAGIC   12582912.0f

static inline float fm_from_bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t fm_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

static inline float fast_exp_pm(float x, float *neg)
{
    if (x < -86.0f) x = -86.0f;
    if (x > 86.0f) x = 86.0f;
    float t = fmaf(x, FM_LOG2E, FM_MAGIC);
    float k = t - FM_MAGIC;
    float r = fmaf(-k, FM_LN2_HI, x);
    r = fmaf(-k, FM_LN2_LO, r);
    float r2 = r * r;
    float r4 = r2 * r2;
    float A = fmaf(0x1.555556p-5f, r2, 0.5f);
    A = fmaf(0x1.6c16c2p-10f, r4, A);
    float B = fmaf(0x1.111112p-7f, r2, 0x1.555556p-3f);
    B = fmaf(0x1.a01a02p-13f, r4, B);
    uint32_t ks = fm_bits(t) << 23;
    float sp = fmaf(r2, fmaf(r, B, A), r);
    float sn = fmaf(r2, fmaf(-r, B, A), -r);
    *neg = fm_from_bits(fm_bits(1.0f + sn) - ks);
    return fm_from_bits(fm_bits(1.0f + sp) + ks);
}

static inline float fast_expf(float x)
{
    float unused;
    return fast_exp_pm(x, &unused);
}

/* x: finite, positive, normal */
static inline float fast_logf(float x)
{
    uint32_t ix = fm_bits(x);
    int32_t e = (int32_t)(ix - 0x3f3504f3u) >> 23;
    float m = fm_from_bits(ix - ((uint32_t)e << 23));
    float f = m - 1.0f;
    float f2 = f * f;
    float f4 = f2 * f2;
    float p01 = fmaf(-0x1.000364p-2f, f, 0x1.5557c8p-2f);
    float p23 = fmaf(-0x1.537288p-3f, f, 0x1.989ae0p-3f);
    float p45 = fmaf(-0x1.2534cap-3f, f, 0x1.33a014p-3f);
    float q0 = fmaf(p23, f2, p01);
    float q1 = fmaf(0x1.5b4e82p-4f, f2, p45);
    float R = fmaf(q1, f4, q0);
    float f3 = f2 * f;
    float lp = fmaf(f3, R, fmaf(-0.5f, f2, f));
    float ef = (float)e;
    return fmaf(ef, FM_LN2_HI, fmaf(ef, FM_LN2_LO, lp));
}

static inline float fast_hypotf(float a, float b) { return sqrtf(fmaf(a, a, b * b)); }

#endif
