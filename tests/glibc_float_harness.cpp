// Host build of cimba_b200/csrc/glibc_float.cuh for tests/test_awacs_math.py (device intrinsics -> IEEE meanings).
// Usage: harness <samples> [exhaustive]   - "exhaustive" walks every float with |x| < 88 for expf and every float in
// [2^-31, 2^31] for powf(x, 4.0f) on 8 threads (about 15 s).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline long long __double_as_longlong(double d) { long long i; std::memcpy(&i, &d, 8); return i; }
static inline double __longlong_as_double(long long i) { double d; std::memcpy(&d, &i, 8); return d; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }

#include "../cimba_b200/csrc/glibc_float.cuh"

static inline uint32_t bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static uint64_t state = 88172645463325252ull;
static inline uint64_t next64() { state ^= state << 13; state ^= state >> 7; state ^= state << 17; return state; }
static inline double unit() { return (double)(int64_t)next64() / 9.3e18; }

int main(int argc, char **argv)
{
    const long n = argc > 1 ? std::atol(argv[1]) : 2000000;
    const bool exhaustive = argc > 2;
    unsigned long bad_exp = 0, bad_pow = 0, bad_pow4 = 0;
    for (long i = 0; i < n; i++) {
        const float x = (float)(unit() * ((i & 1) ? 87.9 : 6.0));
        bad_exp += bits(expf(x)) != bits(cimba_b200::glibc_expf(x));
        const float b = (float)(std::fabs(unit()) * ((i & 2) ? 3.0e5 : 2.0) + 1e-6), e = (float)(unit() * 6.0);
        bad_pow += bits(powf(b, e)) != bits(cimba_b200::glibc_powf(b, e));
        bad_pow4 += bits(powf(b, 4.0f)) != bits(cimba_b200::glibc_powf(b, 4.0f));
    }
    unsigned long ex_exp = 0, ex_pow4 = 0;
    if (exhaustive) {
        std::vector<unsigned long> part(16, 0ul);
        std::vector<std::thread> pool;
        for (int t = 0; t < 8; t++) {
            pool.emplace_back([t, &part]() {
                for (uint64_t u = (uint64_t)t; u < (1ull << 32); u += 8) {
                    float x; const uint32_t w = (uint32_t)u; std::memcpy(&x, &w, 4);
                    // the fast paths only: |x| < 88 for expf, 2^-31 <= x <= 2^31 for powf(x, 4) (beyond them the header
                    // falls back to the double routine, and the results are subnormal, zero or infinite)
                    if (((w >> 20) & 0x7ffu) < 0x42bu) part[t] += bits(expf(x)) != bits(cimba_b200::glibc_expf(x));
                    if (w >= 0x30000000u && w <= 0x4f000000u) part[8 + t] += bits(powf(x, 4.0f)) != bits(cimba_b200::glibc_powf(x, 4.0f));
                }
            });
        }
        for (auto &th : pool) th.join();
        for (int t = 0; t < 8; t++) { ex_exp += part[t]; ex_pow4 += part[8 + t]; }
    }
    std::printf("{\"n\": %ld, \"expf\": %lu, \"powf\": %lu, \"powf4\": %lu, \"exhaustive\": %s, \"expf_all_floats\": %lu, \"powf4_all_floats\": %lu}\n",
                n, bad_exp, bad_pow, bad_pow4, exhaustive ? "true" : "false", ex_exp, ex_pow4);
    return 0;
}
