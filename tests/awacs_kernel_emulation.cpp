// tests/awacs_kernel_emulation.cpp - TEST INFRASTRUCTURE: runs the SOURCE TEXT of awacs_kernel
// (cimba_b200/csrc/awacs_model.cuh) on the CPU, one warp = 32 cooperatively scheduled lanes in one OS thread.
//
// Why: the kernel's control flow (target state machine, the three passes of a radar tick, the event selection)
// can be compared with the plain-C oracle for simulated times a GPU test would make slow on the oracle's side,
// and without a GPU at all.  What it is NOT: a product path (it is built by a test, under tests/), a performance
// model, or a check of the device's arithmetic (IEEE operations and the generator come from the host here).
//
// How: every lane is a ucontext coroutine; a warp-level primitive (__ballot_sync, __shfl_sync, __any_sync,
// __reduce_*_sync, __syncwarp, __syncthreads) is a rendezvous: a lane deposits its value, yields round-robin until all
// 32 have arrived, and reads the result.  The CUDA qualifiers and the arithmetic intrinsics the header uses are given
// their obvious host meanings; the pieces awacs_model.cuh takes from engine.cuh / hold_deep.cuh / rng.cuh (the
// generator, warp_first, the status bits) are supplied here on top of the oracle's generator, which the GPU tests
// of every other model already hold the device generator to.
//
// Build (tests/test_awacs_kernel_emulation.py):
//   g++ -std=c++17 -O2 -ffp-contract=off awacs_kernel_emulation.cpp -o emu -L../oracle -loracle_port -Wl,-rpath,...
// Usage: emu <width_nm> <height_nm> <seconds> <trial index> -> one JSON line comparing the emulated kernel with the oracle.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <ucontext.h>
#include <unistd.h>

// ------------------------------------------------------------------ the oracle (checker and generator source)
extern "C" {
typedef struct { uint64_t a, b, c, d; } port_rng;
uint64_t port_fmix64(uint64_t seed, uint64_t nonce);
void     port_rng_init(port_rng *r, uint64_t seed);
double   port_random(port_rng *r);
double   port_uniform(port_rng *r, double lo, double hi);
double   port_exponential(port_rng *r, double mean);
double   port_erlang(port_rng *r, unsigned k, double m);
unsigned port_bernoulli(port_rng *r, double p);
typedef struct {
    uint64_t events; double t_end; uint32_t num_found; uint32_t tds_count[6]; uint32_t mode_count[4]; uint32_t pad;
    double sum_x, sum_y;
} port_awacs_out;
void port_awacs_grid(float width_nm, float height_nm, uint32_t *cols, uint32_t *rows);
int  port_awacs_terrain_mt(uint64_t seed, float width_nm, float height_nm, float ref_lat, float ref_lon,
                           float *map, float *geom, int *blueprint_out, int threads);
int  port_awacs_trial(uint64_t seed, double duration_h, const float *map, uint32_t cols, uint32_t rows,
                      const float *geom, uint64_t trace_cap, uint64_t *trace_key, double *trace_time,
                      port_awacs_out *out, float *xs, float *ys, int *modes, int *tdss, int *dets);
}

struct Dim3 { unsigned x, y, z; };
static Dim3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {128, 1, 1};

// ------------------------------------------------------------------ 32 lanes as coroutines
namespace emu {
constexpr int LANES = 32;
ucontext_t main_ctx, lane_ctx[LANES];
bool finished[LANES];
int current = 0;
unsigned long long deposit[LANES];
int arrived = 0;
unsigned long generation = 0;

void yield_to_next()
{
    const int from = current;
    int to = from;
    for (int step = 1; step <= LANES; step++) {
        const int cand = (from + step) % LANES;
        if (!finished[cand]) { to = cand; break; }
    }
    if (to == from) return;
    current = to;
    swapcontext(&lane_ctx[from], &lane_ctx[to]);
    threadIdx.x = (unsigned)from;       // back on this lane
}

// all 32 lanes meet
void meet()
{
    const unsigned long gen = generation;
    if (++arrived == LANES) {
        arrived = 0;
        generation++;
    }
    else {
        while (generation == gen) yield_to_next();
    }
}
// meet with a value each: afterwards deposit[] holds every lane's value ...
void rendezvous(unsigned long long mine)
{
    deposit[current] = mine;
    meet();
}
// ... until a second meeting lets the lanes go on (nobody overwrites deposit[] before everybody has read it)
void release() { meet(); }
}  // namespace emu

// ------------------------------------------------------------------ CUDA vocabulary on the host
#define __global__
#define __device__
#define __forceinline__ inline
#define __noinline__
#define __shared__ static
#define __launch_bounds__(x)
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline int __double2int_rz(double d) { return (int)d; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline long long __double_as_longlong(double d) { long long i; std::memcpy(&i, &d, 8); return i; }
static inline double __longlong_as_double(long long i) { double d; std::memcpy(&d, &i, 8); return d; }
static inline float __ldg(const float *p) { return *p; }
static inline int __popc(unsigned m) { return __builtin_popcount(m); }
static inline int __ffs(unsigned m) { return __builtin_ffs((int)m); }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }

static inline void __syncwarp() { emu::meet(); }
static inline void __syncthreads() { emu::meet(); }
static inline unsigned __ballot_sync(unsigned, bool p)
{
    emu::rendezvous(p ? 1ull : 0ull);
    unsigned m = 0u;
    for (int l = 0; l < emu::LANES; l++) m |= emu::deposit[l] ? (1u << l) : 0u;
    emu::release();
    return m;
}
static inline bool __any_sync(unsigned mask, bool p) { return __ballot_sync(mask, p) != 0u; }
template <class T> static inline T __shfl_sync(unsigned, T v, unsigned src)
{
    unsigned long long bits = 0ull;
    std::memcpy(&bits, &v, sizeof(T));
    emu::rendezvous(bits);
    const unsigned long long got = emu::deposit[src & 31u];
    emu::release();
    T out;
    std::memcpy(&out, &got, sizeof(T));
    return out;
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask)
{
    return __shfl_sync(mask, v, (unsigned)(emu::current ^ lane_mask));
}
static inline unsigned __reduce_add_sync(unsigned, unsigned v)
{
    emu::rendezvous(v);
    unsigned s = 0u;
    for (int l = 0; l < emu::LANES; l++) s += (unsigned)emu::deposit[l];
    emu::release();
    return s;
}
static inline unsigned __reduce_min_sync(unsigned, unsigned v)
{
    emu::rendezvous(v);
    unsigned s = 0xffffffffu;
    for (int l = 0; l < emu::LANES; l++) s = (unsigned)emu::deposit[l] < s ? (unsigned)emu::deposit[l] : s;
    emu::release();
    return s;
}

// ------------------------------------------------------------------ what awacs_model.cuh takes from the engine headers
namespace cimba_b200 {
enum : uint32_t { TRIAL_OK = 0u, TRIAL_ERR_KEY_OVERFLOW = 4u };                 // engine.cuh
static inline uint64_t fmix64(uint64_t seed, uint64_t nonce) { return port_fmix64(seed, nonce); }   // rng.cuh
struct ZigHot { double exp_x[256]; double nor_x[256]; };
static inline void stage_zig_hot(ZigHot &, bool) {}
struct Sfc64 {                                                                  // the calls awacs_kernel makes
    port_rng r;
    void seed(uint64_t s) { port_rng_init(&r, s); }
    double uniform01() { return port_random(&r); }
    double uniform(double lo, double hi) { return port_uniform(&r, lo, hi); }
    unsigned bernoulli(double p) { return port_bernoulli(&r, p); }
    double exponential(const ZigHot &, double mean) { return port_exponential(&r, mean); }
    double erlang(const ZigHot &, unsigned k, double m) { return port_erlang(&r, k, m); }
};
struct Picked { unsigned lane; unsigned long long t; uint32_t key; };          // hold_deep.cuh, same text
static inline Picked warp_first(unsigned long long ct, uint32_t ck)
{
    constexpr unsigned FULL = 0xffffffffu;
    const uint32_t hi = (uint32_t)(ct >> 32), lo = (uint32_t)ct;
    const uint32_t mhi = __reduce_min_sync(FULL, hi);
    unsigned cand = __ballot_sync(FULL, hi == mhi);
    if (__popc(cand) > 1) {
        const uint32_t mlo = __reduce_min_sync(FULL, hi == mhi ? lo : 0xffffffffu);
        const bool tie = (hi == mhi) & (lo == mlo);
        const uint32_t mkey = __reduce_min_sync(FULL, tie ? ck : 0xffffffffu);
        cand = __ballot_sync(FULL, tie & (ck == mkey));
    }
    Picked p;
    p.lane = __ffs(cand) - 1u;
    p.t = __shfl_sync(FULL, ct, p.lane);
    p.key = __shfl_sync(FULL, ck, p.lane);
    return p;
}
static inline bool goes_before(unsigned long long at, uint32_t ak, unsigned long long bt, uint32_t bk)
{
    return at < bt || (at == bt && ak < bk);
}
}  // namespace cimba_b200

// tma_bulk.cuh on the host.  An mbarrier is its 8 bytes: low word = phases completed, high word = bytes still expected
// in the current phase.  A copy happens when it is issued and counts its bytes down; the phase completes when the
// expected bytes have arrived; a waiter yields to the other lanes until the phase of the parity it names has completed
// (the barrier's current parity differs from it) - the device semantics, minus the asynchrony.
namespace cimba_b200 { namespace tma {
static inline void barrier_init(uint64_t *bar, uint32_t) { *bar = 0u; }
static inline void barrier_init_fence() {}
static inline void barrier_expect(uint64_t *bar, uint32_t bytes) { *bar += (uint64_t)bytes << 32; }
static inline void barrier_wait(uint64_t *bar, uint32_t parity)
{
    while ((uint32_t)(*bar & 1u) == parity) emu::yield_to_next();
}
static inline void load(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    std::memcpy(dst, src, bytes);
    *bar -= (uint64_t)bytes << 32;
    if ((*bar >> 32) == 0u) *bar += 1u;                 // all announced bytes are in: the phase completes
}
static inline void fence_global_to_async() {}
} }
#define __align__(n)

#define AWACS_HOST_EMULATION 1
#include "../cimba_b200/csrc/awacs_model.cuh"

// ------------------------------------------------------------------ the host half of the launch (capi.cu: awacs_orbit)
static cimba_b200::AwacsOrbit orbit_constants()
{
    using namespace cimba_b200;
    const double PI = 3.14159265358979323846;
    const double deg_to_rad = (2.0 * PI / 360.0), nm_to_meters = 1852.0, feet_to_meters = 0.3048, knots_to_ms = (1852.0 / 3600.0);
    const double WGS84_A = 6378137.0, WGS84_F = (1.0 / 298.257223563), WGS84_E2 = (WGS84_F * (2.0 - WGS84_F));
    AwacsOrbit o{};
    o.start_time = 0.0f;
    const float anchor_lat_r = (float)(30.0f * deg_to_rad);
    o.orientation_r = (float)((90.0 - 0.0f) * deg_to_rad);
    o.length_m = (float)(50.0f * nm_to_meters);
    o.turn_radius_m = (float)(10.0f * nm_to_meters);
    o.altitude_m = (float)(310.0f * 100.0 * feet_to_meters);
    o.velocity_ms = (float)(300.0f * knots_to_ms);
    o.turn_dist_m = (float)(PI * o.turn_radius_m);
    o.orbit_dist_m = 2.0f * (o.length_m + o.turn_dist_m);
    o.side = -1.0f;
    const double sin_lat = sinf(anchor_lat_r);
    const double common = 1.0 - (WGS84_E2 * sin_lat * sin_lat);
    const double sqrt_common = sqrt(common);
    const double M = WGS84_A * (1.0 - WGS84_E2) / (common * sqrt_common);
    const double N = WGS84_A / sqrt_common;
    const double roll_mag = atan((o.velocity_ms * o.velocity_ms) / (o.turn_radius_m * 9.80665));
    o.roll_angle_r = (float)(roll_mag * -o.side);
    o.rad_eff = (float)(sqrt(M * N) * (4.0 / 3.0));
    o.cos_o = cos((double)o.orientation_r);
    o.sin_o = sin((double)o.orientation_r);
    return o;
}

static cimba_b200::AwacsArgs g_args;

static void lane_main(int lane)
{
    threadIdx.x = (unsigned)lane;       // read by the kernel before its first rendezvous only
    cimba_b200::awacs_kernel<true>(g_args);
    emu::finished[lane] = true;
    // hand over to a lane that is still running, or back to main when none is
    for (int step = 1; step <= emu::LANES; step++) {
        const int cand = (lane + step) % emu::LANES;
        if (!emu::finished[cand]) {
            emu::current = cand;
            setcontext(&emu::lane_ctx[cand]);
        }
    }
    setcontext(&emu::main_ctx);
}

#include <execinfo.h>
#include <signal.h>
static void on_segv(int)
{
    void *bt[40];
    const int n = backtrace(bt, 40);
    std::fprintf(stderr, "SIGSEGV on lane %d, rendezvous generation %lu\n", emu::current, emu::generation);
    backtrace_symbols_fd(bt, n, 2);
    _exit(3);
}

int main(int argc, char **argv)
{
    using namespace cimba_b200;
    static char altstack[1 << 16];
    stack_t ss{};
    ss.ss_sp = altstack;
    ss.ss_size = sizeof(altstack);
    sigaltstack(&ss, nullptr);
    struct sigaction sa{};
    sa.sa_handler = on_segv;
    sa.sa_flags = SA_ONSTACK;
    sigaction(SIGSEGV, &sa, nullptr);
    const float width = argc > 1 ? (float)atof(argv[1]) : 6.0f, height = argc > 2 ? (float)atof(argv[2]) : 5.0f;
    const int seconds = argc > 3 ? atoi(argv[3]) : 120;
    const uint64_t trial_index = argc > 4 ? (uint64_t)atoll(argv[4]) : 0u;
    const uint64_t MASTER = 0x34F05C64D7AD598Full, cap = 8192u;

    uint32_t cols, rows;
    port_awacs_grid(width, height, &cols, &rows);
    std::vector<float> map((size_t)cols * rows);
    float geom[6];
    port_awacs_terrain_mt(MASTER, width, height, 30.0f, -10.0f, map.data(), geom, nullptr, 4);

    std::vector<unsigned char> state(AWACS_STATE_BYTES);
    std::vector<uint64_t> tkey(cap), counters(8);
    std::vector<double> ttime(cap);
    uint64_t events = 0, found = 0;
    double t_end = 0.0, sum_x = 0.0;
    uint32_t status = 0, maxq = 0;
    g_args = AwacsArgs{};
    g_args.master_seed = MASTER;
    g_args.first_trial = trial_index;
    g_args.num_trials = 1;
    g_args.t_end_s = (double)seconds;
    // the tile-maximum map (capi.cu builds it with aw_tile_max_kernel when a terrain is registered)
    const uint32_t tcols = (cols + AWACS_TILE - 1u) >> AWACS_TILE_SHIFT, trows = (rows + AWACS_TILE - 1u) >> AWACS_TILE_SHIFT;
    std::vector<float> tiles((size_t)tcols * trows, -3.402823466e+38f);
    for (uint32_t r = 0; r < rows; r++) {
        for (uint32_t c = 0; c < cols; c++) {
            float &t = tiles[(size_t)(r >> AWACS_TILE_SHIFT) * tcols + (c >> AWACS_TILE_SHIFT)];
            t = std::fmax(t, map[(size_t)r * cols + c]);
        }
    }
    g_args.ter = AwacsTerrain{map.data(), cols, rows, geom[0], geom[1], geom[2], geom[3], geom[4], geom[5], tiles.data(), tcols, trows};
    g_args.orbit = orbit_constants();
    g_args.state = state.data();
    g_args.events = &events; g_args.objects = &found; g_args.t_end = &t_end; g_args.sum_wait = &sum_x;
    g_args.status = &status; g_args.max_queue = &maxq; g_args.counters = counters.data();
    g_args.trace_cap = cap; g_args.trace_key = tkey.data(); g_args.trace_time = ttime.data();

    std::vector<std::vector<char>> stacks(emu::LANES, std::vector<char>(1 << 20));
    for (int l = 0; l < emu::LANES; l++) {
        getcontext(&emu::lane_ctx[l]);
        emu::lane_ctx[l].uc_stack.ss_sp = stacks[l].data();
        emu::lane_ctx[l].uc_stack.ss_size = stacks[l].size();
        emu::lane_ctx[l].uc_link = &emu::main_ctx;
        makecontext(&emu::lane_ctx[l], (void (*)())lane_main, 1, l);
    }
    emu::current = 0;
    swapcontext(&emu::main_ctx, &emu::lane_ctx[0]);


    // ---- the oracle on the same trial
    std::vector<uint64_t> okey(cap);
    std::vector<double> otime(cap);
    std::vector<float> ox(1000), oy(1000);
    std::vector<int> omode(1000), otds(1000), odet(1000);
    port_awacs_out o;
    port_awacs_trial(port_fmix64(MASTER, trial_index), seconds / 3600.0, map.data(), cols, rows, geom, cap, okey.data(), otime.data(),
                     &o, ox.data(), oy.data(), omode.data(), otds.data(), odet.data());

    AwacsState S(state.data());
    unsigned x_diff = 0, tds_diff = 0, mode_diff = 0, found_diff = 0;
    for (int i = 0; i < 1000; i++) {
        x_diff += std::memcmp(&S.x[i], &ox[i], 4) != 0 || std::memcmp(&S.y[i], &oy[i], 4) != 0;
        tds_diff += (int)((S.flags[i] >> 4) & 7u) != otds[i];
        mode_diff += (int)(S.flags[i] & 3u) != omode[i];
        found_diff += (int)((S.flags[i] >> 8) & 1u) != odet[i];
    }
    uint64_t n = events < o.events ? events : o.events;
    if (n > cap) n = cap;
    long first_trace_diff = -1;
    for (uint64_t k = 0; k < n; k++) {
        if (tkey[k] != okey[k] || ttime[k] != otime[k]) { first_trace_diff = (long)k; break; }
    }
    std::printf("{\"seconds\": %d, \"trial\": %llu, \"events\": [%llu, %llu], \"found\": [%llu, %u], \"t_end_equal\": %s, \"status\": %u, "
                "\"position_diffs\": %u, \"tds_diffs\": %u, \"mode_diffs\": %u, \"found_diffs\": %u, \"first_trace_diff\": %ld, "
                "\"modes\": [%u, %u, %u, %u], \"compared_pops\": %llu, \"cells_read\": %llu}\n",
                seconds, (unsigned long long)trial_index, (unsigned long long)events, (unsigned long long)o.events,
                (unsigned long long)found, o.num_found, t_end == o.t_end ? "true" : "false", status, x_diff, tds_diff, mode_diff,
                found_diff, first_trace_diff, o.mode_count[0], o.mode_count[1], o.mode_count[2], o.mode_count[3], (unsigned long long)n,
                (unsigned long long)counters[7]);
    return 0;
}
