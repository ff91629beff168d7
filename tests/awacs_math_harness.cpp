// Host build of cimba_b200/csrc/awacs_math.cuh for tests/test_awacs_math.py: the device intrinsics the header uses are
// given their IEEE-754 meanings, then every routine is compared with the host's libm bit for bit.
// Build: g++ -std=c++17 -O2 -ffp-contract=off awacs_math_harness.cpp -o harness
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline long long __double_as_longlong(double d) { long long i; std::memcpy(&i, &d, 8); return i; }
static inline double __longlong_as_double(long long i) { double d; std::memcpy(&d, &i, 8); return d; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline int __double2int_rz(double d) { return (int)d; }

#include "../cimba_b200/csrc/awacs_math.cuh"

static uint64_t state = 88172645463325252ull;
static inline uint64_t next64() { state ^= state << 13; state ^= state >> 7; state ^= state << 17; return state; }
static inline double unit() { return (double)(int64_t)next64() / 9.3e18; }             // (-1, 1)
static inline uint32_t bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }

int main(int argc, char **argv)
{
    const long n = argc > 1 ? std::atol(argv[1]) : 2000000;
    unsigned long bad_atan2 = 0, bad_sin = 0, bad_cos = 0, diff_pow = 0, diff_exp = 0;
    for (long i = 0; i < n; i++) {
        float y, x;
        switch (i & 3) {
        case 0: y = (float)(unit() * 2.0e5); x = (float)(unit() * 2.0e5); break;        // (dy, dx) in metres
        case 1: y = (float)(unit() * 1.2e4); x = (float)(std::fabs(unit()) * 4.0e5); break;   // (dz, d_2d)
        case 2: { uint32_t a = (uint32_t)next64(), b = (uint32_t)next64(); std::memcpy(&y, &a, 4); std::memcpy(&x, &b, 4); break; }
        default: y = (float)(unit() * 8.0); x = (float)(unit() * 8.0); break;
        }
        if (std::isfinite(x) && std::isfinite(y) && bits(atan2f(y, x)) != bits(cimba_b200::aw_atan2f(y, x))) bad_atan2++;
        const float a = (float)(unit() * ((i & 1) ? 13.0 : 119.0));                     // the model stays inside +-4 pi
        if (bits(sinf(a)) != bits(cimba_b200::aw_sinf(a))) bad_sin++;
        if (bits(cosf(a)) != bits(cimba_b200::aw_cosf(a))) bad_cos++;
        const float r = (float)(std::fabs(unit()) * 30.0 + 1.0e-3);
        if (bits(powf(r, 4.0f)) != bits(cimba_b200::aw_powf(r, 4.0f))) diff_pow++;
        const float e = (float)(unit() * 40.0);
        if (bits(expf(e)) != bits(cimba_b200::aw_expf(e))) diff_exp++;
    }
    std::printf("{\"n\": %ld, \"atan2f\": %lu, \"sinf\": %lu, \"cosf\": %lu, \"powf_rounded_once\": %lu, \"expf_rounded_once\": %lu}\n",
                n, bad_atan2, bad_sin, bad_cos, diff_pow, diff_exp);
    return 0;
}
