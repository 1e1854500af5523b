// detmath.cuh -- bit-reproducible exp / log / hypot for the stroke-expansion kernel.
//
// libmypaint 1.3.0's brush engine calls expf/exp/log/hypotf from libm (mypaint-brush.c:
// exp_decay, update_states_and_setting_values, count_dabs_to).  CUDA's libdevice versions differ
// from glibc's in the last ulp, and the dab loop is discontinuous in its inputs (one more or one
// fewer dab shifts the RNG stream), so the rasteriser uses its own functions built only from
// IEEE-754 correctly rounded operations (+ - * / sqrt fma, conversions).  A CPU program that
// evaluates the same expression tree gets the same bits -- that is what the parity tests rely on.
//
// Compile every TU that includes this header with  --fmad=false  (no implicit contraction);
// fused operations are written explicitly as fma().
//
// det_exp(x):  k = rint(x*log2e); r = x - k*ln2 (two-step, fma); degree-12 Taylor of exp(r)
//              evaluated as an Estrin tree; scaled by 2^k in two exact steps.  |rel err| < 1e-15.
// det_log(x):  x = m*2^e, m in (sqrt(.5), sqrt(2)]; t = (m-1)/(m+1); log m = t*q(t^2), q of
//              degree 8 in t^2 with c_k = 2/(2k+1); result = e*ln2 + t*q.       |rel err| < 1e-15.
// det_hypotf(a,b) = (float) sqrt((double)a*a + (double)b*b)   (both products are exact).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define DM_HD __host__ __device__ __forceinline__
#else
#define DM_HD inline
#endif

namespace detmath {

DM_HD double bits_to_double(uint64_t u)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d;
    memcpy(&d, &u, 8);
    return d;
#endif
}
DM_HD uint64_t double_to_bits(double d)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u;
    memcpy(&u, &d, 8);
    return u;
#endif
}
DM_HD double pow2i(int k) { return bits_to_double((uint64_t)(k + 1023) << 52); }

DM_HD double det_exp(double x)
{
    // |x| <= 700 (false for NaN) is the only case the rasteriser ever sees: one comparison instead of three, and the
    // result 2^k * p is a normal number, so ONE exact scaling gives the same bits as the two-step one below
    const bool small = fabs(x) <= 700.0;
    if (!small) {
        if (x != x) return x;
        if (x > 709.0) return bits_to_double(0x7ff0000000000000ULL);
        if (x < -745.0) return 0.0;
    }
    const double kd = rint(x * 0x1.71547652b82fep+0);
    double r = fma(-kd, 0x1.62e42fee00000p-1, x);
    r = fma(-kd, 0x1.a39ef35793c76p-33, r);
    const double r2 = r * r;
    const double r4 = r2 * r2;
    const double r8 = r4 * r4;
    const double a01 = fma(1.0, r, 1.0);
    const double a23 = fma(0x1.5555555555555p-3, r, 0x1.0000000000000p-1);
    const double a45 = fma(0x1.1111111111111p-7, r, 0x1.5555555555555p-5);
    const double a67 = fma(0x1.a01a01a01a01ap-13, r, 0x1.6c16c16c16c17p-10);
    const double a89 = fma(0x1.71de3a556c734p-19, r, 0x1.a01a01a01a01ap-16);
    const double aab = fma(0x1.ae64567f544e4p-26, r, 0x1.27e4fb7789f5cp-22);
    const double b0 = fma(a23, r2, a01);
    const double b1 = fma(a67, r2, a45);
    double b2 = fma(aab, r2, a89);
    b2 = fma(0x1.1eed8eff8d898p-29, r4, b2);
    double p = fma(b1, r4, b0);
    p = fma(b2, r8, p);
    const int ki = (int)kd;
    if (small) return p * pow2i(ki);
    const int k1 = ki / 2;
    const int k2 = ki - k1;
    return p * pow2i(k1) * pow2i(k2);
}

// x: finite, positive, normal
DM_HD double det_log(double x)
{
    const uint64_t b = double_to_bits(x);
    int e = (int)((b >> 52) & 0x7ff) - 1023;
    double m = bits_to_double((b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
    if (m > 0x1.6a09e667f3bcdp+0) {
        m = m * 0.5;
        e += 1;
    }
    const double t = (m - 1.0) / (m + 1.0);
    const double w = t * t;
    const double w2 = w * w;
    const double w4 = w2 * w2;
    const double w8 = w4 * w4;
    const double a01 = fma(0x1.5555555555555p-1, w, 0x1.0000000000000p+1);
    const double a23 = fma(0x1.2492492492492p-2, w, 0x1.999999999999ap-2);
    const double a45 = fma(0x1.745d1745d1746p-3, w, 0x1.c71c71c71c71cp-3);
    const double a67 = fma(0x1.1111111111111p-3, w, 0x1.3b13b13b13b14p-3);
    const double b0 = fma(a23, w2, a01);
    const double b1 = fma(a67, w2, a45);
    double q = fma(b1, w4, b0);
    q = fma(0x1.e1e1e1e1e1e1ep-4, w8, q);
    const double ed = (double)e;
    return fma(ed, 0x1.62e42fee00000p-1, fma(t, q, ed * 0x1.a39ef35793c76p-33));
}

DM_HD float det_expf(float x) { return (float)det_exp((double)x); }

DM_HD float det_hypotf(float a, float b)
{
    const double da = (double)a, db = (double)b;
    return (float)sqrt(da * da + db * db);
}

}  // namespace detmath
