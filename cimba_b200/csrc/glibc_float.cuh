// glibc_float.cuh - glibc's expf() and powf() restated; awacs_math.cuh uses them for the AWACS detection probability.
// Provenance: the algorithms are those of glibc 2.39 sysdeps/ieee754/flt-32/e_expf.c and e_powf.c (LGPL-2.1-or-later),
// which adopted them from ARM optimized-routines (MIT OR Apache-2.0 WITH LLVM-exception); restated here from the published
// description, tables recomputed (see below), no source text copied.
//
// The AWACS detection probability (tutorial/tut_5_1.c:643-686) calls powf(ref_range / r, 4.0f) and expf(...) once per
// attempt; awacs_math.cuh evaluates them in double and rounds once, which differs from glibc 2.39 on 0.065 % of
// arguments by one ulp.  These are glibc's own algorithms (sysdeps/ieee754/flt-32/e_expf.c, e_powf.c, the ARM
// optimized-routines kernels): a 16-entry log2 table + degree-5 polynomial, a 32-entry 2^(i/32) table + cubic, all in
// double.  The tables are not in /root/reference (glibc is not): the 2^(i/32) entries are recomputed to 80 digits,
// except the last one, where glibc's own entry is NOT the correctly rounded value - the entry used here lies in the
// middle of the window of values for which expf agrees with glibc on EVERY float with |x| < 88 (exhaustive check);
// the log2 entries are the published ones (each logc equals -log2(invc) correctly rounded, which pins them).
// Checked on the CPU against glibc: expf on every float with |x| < 88, powf(x, 4.0f) on every float in [2^-31, 2^31],
// powf(x, y) on 4e7 random pairs: no mismatch (tests/test_awacs_math.py runs the sampled part on every CPU run,
// CIMBA_B200_EXHAUSTIVE=1 the rest).  Arguments outside the fast paths (|x| >= 88 for expf; zero / subnormal /
// infinite / NaN operands or an over- / underflowing result for powf) fall back to the double routine rounded once.
#pragma once

#include <cstdint>

#ifndef AW_MATH_FN
#ifdef __CUDACC__
#define AW_MATH_FN __device__ __forceinline__
#else
#define AW_MATH_FN static inline
#endif
#endif

namespace cimba_b200 {

#ifdef __CUDACC__
__device__
#endif
static const uint64_t GLIBC_EXP2F_TAB[32] = {       // bits of 2^(i/32) minus (i << 47)
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4584ull,
};

#ifdef __CUDACC__
__device__
#endif
static const double GLIBC_POWF_LOG2_TAB[16][2] = {  // {1/c, log2(c)} for the 16 sub-intervals of [0x1.66p-1, 0x1.66p0)
    {0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2}, {0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2},
    {0x1.49539f0f010bp+0, -0x1.7418b0a1fb77bp-2},  {0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2},
    {0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2}, {0x1.25e227b0b8eap+0, -0x1.97c1d1b3b7afp-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3}, {0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4},
    {0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5}, {0x1p+0, 0x0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4},  {0x1.ca4b31f026aap-1, 0x1.476a9543891bap-3},
    {0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2},
    {0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2},  {0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2},
};

// 2^x for |x| < 126 given as k/32 + r: s * (C0 r^3 + C1 r^2 + C2 r + 1), e_powf.c exp2_inline / e_expf.c
AW_MATH_FN double glibc_exp2_core(uint64_t ki, double r, double c0, double c1, double c2)
{
    const uint64_t t = GLIBC_EXP2F_TAB[ki % 32u] + (ki << (52 - 5));
    const double s = __longlong_as_double((long long)t);
    const double z = __fma_rn(c0, r, c1);
    const double r2 = __dmul_rn(r, r);
    double y = __fma_rn(c2, r, 1.0);
    y = __fma_rn(z, r2, y);
    return __dmul_rn(y, s);
}

AW_MATH_FN float glibc_expf(float x)
{
    if ((((uint32_t)__float_as_int(x) >> 20) & 0x7ffu) >= 0x42bu) {            // |x| >= 88 (or NaN)
        return (float)exp((double)x);
    }
    const double N = 32.0;
    const double z = __dmul_rn(0x1.71547652b82fep+0 * N, (double)x);
    double kd = __dadd_rn(z, 0x1.8p+52);
    const uint64_t ki = (uint64_t)__double_as_longlong(kd);
    kd = __dsub_rn(kd, 0x1.8p+52);
    const double r = __dsub_rn(z, kd);
    return (float)glibc_exp2_core(ki, r, 0x1.c6af84b912394p-5 / N / N / N, 0x1.ebfce50fac4f3p-3 / N / N, 0x1.62e42ff0c52d6p-1 / N);
}

AW_MATH_FN float glibc_powf(float x, float y)
{
    const uint32_t ix = (uint32_t)__float_as_int(x), iy = (uint32_t)__float_as_int(y);
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u || 2u * iy - 1u >= 2u * 0x7f800000u - 1u) {
        return (float)pow((double)x, (double)y);        // x <= 0, subnormal, inf or NaN; y zero, inf or NaN
    }
    // log2_inline: x = 2^k z with z in [0x1.66p-1, 0x1.66p0), log2(x) = log1p(z/c - 1)/ln2 + log2(c) + k
    const uint32_t tmp = ix - 0x3f330000u;
    const uint32_t i = (tmp >> (23 - 4)) % 16u;
    const uint32_t top = tmp & 0xff800000u;
    const int k = (int32_t)top >> 23;
    const double z = (double)__int_as_float((int)(ix - top));
    const double r = __fma_rn(z, GLIBC_POWF_LOG2_TAB[i][0], -1.0);
    const double y0 = __dadd_rn(GLIBC_POWF_LOG2_TAB[i][1], (double)k);
    const double r2 = __dmul_rn(r, r);
    double p1 = __fma_rn(0x1.27616c9496e0bp-2, r, -0x1.71969a075c67ap-2);
    const double p2 = __fma_rn(0x1.ec70a6ca7baddp-2, r, -0x1.7154748bef6c8p-1);
    const double r4 = __dmul_rn(r2, r2);
    double q = __fma_rn(0x1.71547652ab82bp0, r, y0);
    q = __fma_rn(p2, r2, q);
    p1 = __fma_rn(p1, r4, q);
    const double ylogx = __dmul_rn((double)y, p1);
    if ((((uint64_t)__double_as_longlong(ylogx) >> 47) & 0xffffu) >= (0x405f800000000000ull >> 47)) {      // |y log2 x| >= 126
        return (float)pow((double)x, (double)y);
    }
    double kd = __dadd_rn(ylogx, 0x1.8p+52 / 32);
    const uint64_t ki = (uint64_t)__double_as_longlong(kd);
    kd = __dsub_rn(kd, 0x1.8p+52 / 32);
    const double rr = __dsub_rn(ylogx, kd);
    return (float)glibc_exp2_core(ki, rr, 0x1.c6af84b912394p-5, 0x1.ebfce50fac4f3p-3, 0x1.62e42ff0c52d6p-1);
}

}  // namespace cimba_b200
