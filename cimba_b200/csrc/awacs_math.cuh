// awacs_math.cuh - the float32 libm calls of the AWACS model (awacs_model.cuh), kept apart so that the SAME source
// text can be compiled for the host: tests/test_awacs_math.py builds it with g++ (the device intrinsics below mapped
// to their IEEE meanings) and compares every routine with glibc, bit for bit, on millions of arguments.
#pragma once

#include <cstdint>

#ifdef __CUDACC__
#define AW_MATH_FN __device__ __forceinline__
#else
#define AW_MATH_FN static inline
#endif

namespace cimba_b200 {

// ---- the float32 libm calls of the model.  The oracle is the reference linked against glibc 2.39, so these restate
// that library's published algorithms (not part of /root/reference): atan2f / atanf = the fdlibm float routines
// (sysdeps/ieee754/flt-32/e_atan2f.c, s_atanf.c: argument reduction to four intervals + an 11-term odd polynomial in
// float arithmetic), sinf / cosf = the ARM optimized-routines kernel glibc adopted in 2.28 (sysdeps/ieee754/flt-32/
// s_sincosf.h: quadrant reduction and two degree-7/8 polynomials evaluated in double, here with the multiply-adds
// fused as glibc's x86-64 FMA build does).  Checked on the CPU against glibc itself, bit for bit: 6e7 random
// argument pairs for atan2f, 4e7 arguments each for sinf and cosf, no mismatch.  powf and expf are glibc's table-driven
// float routines restated in glibc_float.cuh (exhaustively checked there); -DAWACS_ROUNDED_ONCE_FLOAT selects the
// round-1 form instead (double results rounded once: differs from glibc in rare last places).
AW_MATH_FN float aw_atanf(float x)
{
    const int32_t hx = __float_as_int(x), ix = hx & 0x7fffffff;
    int id;
    float hi = 0.0f, lo = 0.0f;
    if (ix >= 0x4c000000) {                             // |x| >= 2^25
        if (ix > 0x7f800000) return x + x;
        return hx > 0 ? 1.5707962513e+00f + 7.5497894159e-08f : -1.5707962513e+00f - 7.5497894159e-08f;
    }
    if (ix < 0x3ee00000) {                              // |x| < 7/16
        if (ix < 0x31000000) return x;
        id = -1;
    }
    else {
        x = fabsf(x);
        if (ix < 0x3f980000) {
            if (ix < 0x3f300000) { id = 0; hi = 4.6364760399e-01f; lo = 5.0121582440e-09f; x = __fdiv_rn(2.0f * x - 1.0f, 2.0f + x); }
            else                 { id = 1; hi = 7.8539812565e-01f; lo = 3.7748947079e-08f; x = __fdiv_rn(x - 1.0f, x + 1.0f); }
        }
        else {
            if (ix < 0x401c0000) { id = 2; hi = 9.8279368877e-01f; lo = 3.4473217170e-08f; x = __fdiv_rn(x - 1.5f, 1.0f + 1.5f * x); }
            else                 { id = 3; hi = 1.5707962513e+00f; lo = 7.5497894159e-08f; x = __fdiv_rn(-1.0f, x); }
        }
    }
    const float z = x * x, w = z * z;                   // -fmad=false: every product and sum below rounds separately
    const float s1 = z * (3.3333334327e-01f + w * (1.4285714924e-01f + w * (9.0908870101e-02f + w * (6.6610731184e-02f +
                     w * (4.9768779427e-02f + w * 1.6285819933e-02f)))));
    const float s2 = w * (-2.0000000298e-01f + w * (-1.1111110449e-01f + w * (-7.6918758452e-02f + w * (-5.8335702866e-02f +
                     w * -3.6531571299e-02f))));
    if (id < 0) return x - x * (s1 + s2);
    const float r = hi - ((x * (s1 + s2) - lo) - x);
    return hx < 0 ? -r : r;
}

AW_MATH_FN float aw_atan2f(float y, float x)
{
    const float pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    const int32_t hx = __float_as_int(x), hy = __float_as_int(y), ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
    if (ix == 0x7f800000 || iy == 0x7f800000) return (float)atan2((double)y, (double)x);    // never in this model
    if (hx == 0x3f800000) return aw_atanf(y);
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
    if (iy == 0) return m < 2 ? y : (m == 2 ? pi : -pi);
    if (ix == 0) return hy < 0 ? -pi_o_2 : pi_o_2;
    float z;
    const int k = (iy - ix) >> 23;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
    else if (hx < 0 && k < -60) z = 0.0f;
    else z = aw_atanf(fabsf(__fdiv_rn(y, x)));
    if (m == 0) return z;
    if (m == 1) return -z;
    if (m == 2) return pi - (z - pi_lo);
    return (z - pi_lo) - pi;
}

template <bool WANT_COS>
AW_MATH_FN float aw_sincosf(float y)
{
    const uint32_t top = ((uint32_t)__float_as_int(y) >> 20) & 0x7ffu;
    if (top >= 0x42fu) {                                // |y| >= 120: outside the fast path; never in this model
        return WANT_COS ? (float)cos((double)y) : (float)sin((double)y);
    }
    double x = (double)y;
    int n = 0;
    double sgn = 1.0;
    if (top >= 0x3f4u) {                                // |y| >= pi/4 (by exponent + 3 mantissa bits): reduce
        const double r = __dmul_rn(x, 0x1.45F306DC9C883p+23);
        n = (__double2int_rz(r) + 0x800000) >> 24;
        x = __fma_rn(-(double)n, 0x1.921FB54442D18p0, x);
        sgn = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    }
    else if (top < 0x398u) {                            // |y| < 2^-12
        return WANT_COS ? 1.0f : y;
    }
    const double flip = (n & 2) ? -1.0 : 1.0;           // the second table negates the cosine coefficients
    const double x2 = __dmul_rn(x, x);
    x = __dmul_rn(x, sgn);
    if (((n ^ (WANT_COS ? 1 : 0)) & 1) == 0) {
        const double x3 = __dmul_rn(x, x2);
        const double s1 = __fma_rn(x2, -0x1.994eb3774cf24p-13, 0x1.1107605230bc4p-7);
        const double x7 = __dmul_rn(x3, x2);
        const double s = __fma_rn(x3, -0x1.555545995a603p-3, x);
        return (float)__fma_rn(x7, s1, s);
    }
    const double x4 = __dmul_rn(x2, x2);
    const double c2 = __fma_rn(x2, flip * 0x1.99343027bf8c3p-16, flip * -0x1.6c087e89a359dp-10);
    const double c1 = __fma_rn(x2, flip * -0x1.ffffffd0c621cp-2, flip * 0x1p0);
    const double x6 = __dmul_rn(x4, x2);
    const double c = __fma_rn(x4, flip * 0x1.55553e1068f19p-5, c1);
    return (float)__fma_rn(x6, c2, c);
}
AW_MATH_FN float aw_sinf(float x) { return aw_sincosf<false>(x); }
AW_MATH_FN float aw_cosf(float x) { return aw_sincosf<true>(x); }
#ifndef AWACS_ROUNDED_ONCE_FLOAT     // the shipped build: glibc's powf / expf restated as well (csrc/glibc_float.cuh), so every
                                     // float32 libm call of the model gives glibc's bits and the trial matches the oracle by construction
}  // namespace cimba_b200
#include "glibc_float.cuh"
namespace cimba_b200 {
AW_MATH_FN float aw_powf(float a, float b) { return glibc_powf(a, b); }
AW_MATH_FN float aw_expf(float x) { return glibc_expf(x); }
#else                                // round-1 form, kept for A/B (scripts/build_variant.py ... -DAWACS_ROUNDED_ONCE_FLOAT)
AW_MATH_FN float aw_powf(float a, float b) { return (float)pow((double)a, (double)b); }
AW_MATH_FN float aw_expf(float x) { return (float)exp((double)x); }
#endif

}  // namespace cimba_b200
