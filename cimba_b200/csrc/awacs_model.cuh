// awacs_model.cuh - the reference's AWACS model (tutorial/tut_5_1.c, BASELINE config 5), one trial per WARP.
//
// 1000 ground targets cycle hiding -> staging -> firing -> driving with exponential / Erlang dwell times
// (tut_5_1.c:385-446); a radar process ticks every second and runs a five-stage detection chain over all of
// them in float32 (:944-1036, :548-689) against a terrain map SHARED by every trial (one float per arc-second,
// 14.4 GB at the tutorial's 1000 x 1000 nm - it lives in HBM once, read-only); the platform flies a racetrack
// orbit given in closed form (:816-868); a progress-bar process holds 100 times (:1118-1137); an end event
// stops everybody (:1100-1112).  oracle/port/awacs_port.c is the same model in plain C.
//
// Mapping onto the warp:
//   * every lane carries the same generator state and executes every draw (the stream is per trial); the 32
//     lanes differ only in which target they look at;
//   * the event list: each target owns exactly one pending event, kept in its row of the per-trial state block
//     (wake time + key); the radar, the end event and the progress bar own one each in registers.  The earliest
//     target event is cached and found again by a 32-lane scan + REDUX only after a target event ran (about once
//     per simulated second) - pop order is the total order (time, key) whatever the container;
//   * one radar tick = three passes.  A: lanes take targets 32 at a time - dead reckoning, geometry, the three
//     cheap tests (swept sector, horizon, nadir hole); survivors are appended to a list in target order.
//     B: each survivor's line of sight is marched by the WHOLE warp (256 steps per round, eight consecutive
//     ones per lane, one ballot per round): the reference's loop returns at the first step under the terrain,
//     i.e. "any step is", which is order-free; steps are half a cell apart, so a lane reads each cell once.
//     C: the unshielded survivors, in target order, compute their detection probability in parallel and then
//     take their cmb_random_bernoulli draws one after the other - the only part that is serial by definition.
//
// Parity: sqrtf, division, roundf, fmod and the double-precision orbit are exact on both sides; atan2f, sinf and
// cosf restate glibc's own algorithms (below; bit-identical to glibc on every argument tried); powf and expf are
// evaluated in double and rounded once, which agrees with glibc's float routines except in rare last places.  A
// target passes or fails a test differently only if such a last-place difference straddles the threshold, but when
// it does, the trial's one random stream shifts and everything after differs: this model matches the oracle exactly
// until then and statistically after.
#pragma once

#include "awacs_math.cuh"
#ifndef AWACS_HOST_EMULATION      // tests/awacs_kernel_emulation.cpp supplies these pieces itself
#include "engine.cuh"
#include "hold_deep.cuh"
#include "rng.cuh"
#include "tma_bulk.cuh"
#endif

namespace cimba_b200 {

constexpr int AWACS_BLOCK = 128;                // 4 trials per CTA
constexpr int AWACS_TARGETS = 1000;             // NUM_TARGETS, tut_5_1.c:35
constexpr int AWACS_STRIDE = 1024;              // rows per column of the state block
#ifndef AWACS_CHUNK
#define AWACS_CHUNK 4                           // consecutive line-of-sight steps per lane and round: 128-step stretches
                                                // (measured on B200, 4096 trials x 300 s: 8 -> 334 ms, 4 -> 281 ms, 2 -> 314 ms)
#endif
// per-trial state block in HBM/L2, structure of arrays, AWACS_STRIDE entries each:
//   float x, y, alt, dir, vel, time_s, rcs_now; uint32 flags; uint32 wake_key; double wake_t
constexpr size_t AWACS_STATE_BYTES = (size_t)AWACS_STRIDE * (7 * 4 + 4 + 4 + 8);
// pass A streams the six columns it reads (x, y, alt, dir, vel, time_s) through shared memory in tiles of
// AWACS_TILE_TARGETS rows, two tiles in flight per warp, moved by the TMA bulk-copy engine (tma_bulk.cuh)
constexpr int AWACS_TILE_TARGETS = 128;
constexpr int AWACS_TILES = AWACS_STRIDE / AWACS_TILE_TARGETS;      // 8 tiles cover the 1000 targets (+ 24 idle rows)
constexpr int AWACS_STREAM_COLS = 6;

enum : uint32_t { AW_HIDING = 0, AW_STAGING = 1, AW_FIRING = 2, AW_DRIVING = 3 };
enum : uint32_t { AW_UNDETERMINED = 0, AW_BEYOND_HORIZON, AW_NADIR_HOLE, AW_TERRAIN_SHIELDED, AW_MISSED, AW_DETECTED };
constexpr uint32_t AW_F_MODE = 3u, AW_F_TDS_SHIFT = 4u, AW_F_TDS = 7u << 4, AW_F_FOUND = 1u << 8, AW_F_STARTED = 1u << 9;

struct AwacsTerrain {           // struct terrain, tut_5_1.c:96-108
    const float *map;
    uint32_t cols, rows;
    float x_scale, y_scale, x_min, x_max, y_min, y_max;
    // not in the reference: the highest cell of every AWACS_TILE x AWACS_TILE block of the map ([trows][tcols], built
    // once per registered terrain by aw_tile_max_kernel).  The line-of-sight march asks it which 256-step stretches
    // of a ray can touch the ground at all; see pass B.
    const float *tile_max;
    uint32_t tcols, trows;
};

constexpr uint32_t AWACS_TILE_SHIFT = 7u, AWACS_TILE = 1u << AWACS_TILE_SHIFT;      // 128 x 128 cells: 64 KB of map per entry

struct AwacsOrbit {             // struct racetrack after racetrack_initialize (:724-782), evaluated on the host
    float start_time, orientation_r, length_m, turn_radius_m, altitude_m, velocity_ms;
    float turn_dist_m, orbit_dist_m, side, roll_angle_r, rad_eff;
    double cos_o, sin_o;        // cos / sin of orientation_r (:851-853), host libm
};

struct AwacsArgs {
    uint64_t master_seed, first_trial, num_trials;
    double   t_end_s;           // trial duration in seconds (struct trial.duration * 3600, :1213)
    AwacsTerrain ter;
    AwacsOrbit orbit;
    unsigned char *state;       // [num_trials][AWACS_STATE_BYTES]
    uint64_t *events, *objects;
    double   *t_end, *sum_wait;
    uint32_t *status, *max_queue;
    uint64_t *counters;
    uint64_t  trace_cap;
    uint64_t *trace_key;
    double   *trace_time;
};

// terrain_index + terrain_elevation, tut_5_1.c:314-338
// (a cell index fits 32 bits: the tutorial's 60 000 x 60 000 map has 3.6e9 cells; registration refuses more than 2^32 - 1)
__device__ __forceinline__ uint32_t aw_cell(const AwacsTerrain &t, float x, float y)
{
    const int raw_col = (int)roundf(__fdiv_rn(x, t.x_scale)) + (int)(t.cols / 2u);
    const int raw_row = (int)roundf(__fdiv_rn(y, t.y_scale)) + (int)(t.rows / 2u);
    const uint32_t col = (uint32_t)(raw_col < 0 ? 0 : (raw_col >= (int)t.cols ? (int)t.cols - 1 : raw_col));
    const uint32_t row = (uint32_t)(raw_row < 0 ? 0 : (raw_row >= (int)t.rows ? (int)t.rows - 1 : raw_row));
    return row * t.cols + col;
}

// column / row of terrain_index as two numbers (pass B brackets a stretch of a ray with them)
__device__ __forceinline__ void aw_col_row(const AwacsTerrain &t, float x, float y, uint32_t &col, uint32_t &row)
{
    const int raw_col = (int)roundf(__fdiv_rn(x, t.x_scale)) + (int)(t.cols / 2u);
    const int raw_row = (int)roundf(__fdiv_rn(y, t.y_scale)) + (int)(t.rows / 2u);
    col = (uint32_t)(raw_col < 0 ? 0 : (raw_col >= (int)t.cols ? (int)t.cols - 1 : raw_col));
    row = (uint32_t)(raw_row < 0 ? 0 : (raw_row >= (int)t.rows ? (int)t.rows - 1 : raw_row));
}

__device__ __forceinline__ float aw_elevation(const AwacsTerrain &t, float x, float y)
{
    const int raw_col = (int)roundf(__fdiv_rn(x, t.x_scale)) + (int)(t.cols / 2u);
    const int raw_row = (int)roundf(__fdiv_rn(y, t.y_scale)) + (int)(t.rows / 2u);
    const uint32_t col = (uint32_t)(raw_col < 0 ? 0 : (raw_col >= (int)t.cols ? (int)t.cols - 1 : raw_col));
    const uint32_t row = (uint32_t)(raw_row < 0 ? 0 : (raw_row >= (int)t.rows ? (int)t.rows - 1 : raw_row));
    return __ldg(t.map + (size_t)row * t.cols + col);
}

struct AwacsPlatform { float x, y, dir, rol, alt; };

// platform_state_update, tut_5_1.c:816-868 (double precision, no random numbers)
__device__ __noinline__ void aw_platform_at(AwacsPlatform &st, const AwacsOrbit &o, double t)
{
    const double PI = 3.14159265358979323846;
    const double delta_t = t - (double)o.start_time;
    double d = fmod(delta_t * (double)o.velocity_ms, (double)o.orbit_dist_m);
    if (d < 0) d += (double)o.orbit_dist_m;
    double xl, yl, hdg, roll;
    if (d < (double)o.length_m) {
        xl = d; yl = 0.0; hdg = 0.0; roll = 0.0;
    }
    else if (d < (double)(o.length_m + o.turn_dist_m)) {
        const double phi = (d - (double)o.length_m) / (double)o.turn_radius_m - PI / 2.0;
        xl = (double)o.length_m + (double)o.turn_radius_m * cos(phi);
        yl = (double)(o.side * o.turn_radius_m) * (1.0 + sin(phi));
        hdg = (phi + PI / 2.0) * (double)o.side;
        roll = (double)o.roll_angle_r;
    }
    else if (d < 2.0 * (double)o.length_m + (double)o.turn_dist_m) {
        const double d_seg = d - (double)(o.length_m + o.turn_dist_m);
        xl = (double)o.length_m - d_seg;
        yl = (double)o.side * 2.0 * (double)o.turn_radius_m;
        hdg = PI;
        roll = 0.0;
    }
    else {
        const double phi = (d - (2.0 * (double)o.length_m + (double)o.turn_dist_m)) / (double)o.turn_radius_m + PI / 2.0;
        xl = (double)o.turn_radius_m * cos(phi);
        yl = (double)(o.side * o.turn_radius_m) * (1.0 + sin(phi));
        hdg = PI + (phi - PI / 2.0) * (double)o.side;
        roll = (double)o.roll_angle_r;
    }
    st.x = (float)(xl * o.cos_o - yl * o.sin_o);
    st.y = (float)(xl * o.sin_o + yl * o.cos_o);
    st.dir = (float)fmod(hdg + (double)o.orientation_r + 2.0 * PI, 2.0 * PI);
    st.rol = (float)roll;
    st.alt = o.altitude_m;
}

// one trial's state block, column pointers
struct AwacsState {
    float *x, *y, *alt, *dir, *vel, *time_s, *rcs_now;
    uint32_t *flags, *wake_key;
    double *wake_t;
    __device__ __forceinline__ explicit AwacsState(unsigned char *base)
    {
        float *f = reinterpret_cast<float *>(base);
        x = f; y = f + AWACS_STRIDE; alt = f + 2 * AWACS_STRIDE; dir = f + 3 * AWACS_STRIDE; vel = f + 4 * AWACS_STRIDE;
        time_s = f + 5 * AWACS_STRIDE; rcs_now = f + 6 * AWACS_STRIDE;
        flags = reinterpret_cast<uint32_t *>(f + 7 * AWACS_STRIDE);
        wake_key = flags + AWACS_STRIDE;
        wake_t = reinterpret_cast<double *>(wake_key + AWACS_STRIDE);
    }
};

// tgt->rcs_m2[] and tgt->state_time_s[] as run_trial sets them, tut_5_1.c:1153-1166 with :468-475
__device__ __forceinline__ float aw_rcs(uint32_t mode)
{
    return mode == AW_HIDING ? 5.0f : (mode == AW_STAGING ? 100.0f : (mode == AW_FIRING ? 1000.0f : 50.0f));
}
__device__ __forceinline__ float aw_dwell(uint32_t mode)
{
    return mode == AW_HIDING ? __fmul_rn(3.0f, 3600.0f)
                             : (mode == AW_STAGING ? __fmul_rn(5.0f, 60.0f) : (mode == AW_FIRING ? 30.0f : __fmul_rn(1.0f, 3600.0f)));
}

// the geometry of one target against the platform (tut_5_1.c:987-998) and the three cheap tests;
// returns 0 = not in the swept sector (no change), else the tds it earns, or AW_UNDETERMINED + 16 = survivor
struct AwacsSweep { float prev_dir, width, rad_eff, lo, hi; };

__device__ __forceinline__ uint32_t aw_cheap_tests(const AwacsPlatform &h, const AwacsSweep &sw, float tx, float ty, float ta)
{
    const float TWO_PI_F = __fmul_rn(2.0f, (float)3.14159265358979323846);
    const float dx = __fsub_rn(tx, h.x), dy = __fsub_rn(ty, h.y), dz = __fsub_rn(ta, h.alt);
    const float d_2d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    const float azi = aw_atan2f(dy, dx);
    float rel = __fsub_rn(azi, sw.prev_dir);            // target_is_in_swept_sector, :548-561
    while (rel < 0.0f) rel = __fadd_rn(rel, TWO_PI_F);
    while (rel >= TWO_PI_F) rel = __fsub_rn(rel, TWO_PI_F);
    if (!(rel <= sw.width)) return 0u;
    {                                                   // target_is_beyond_horizon, :566-576
        const float hs = fmaxf(0.0f, h.alt), ht = fmaxf(0.0f, ta);
        const float reach = __fadd_rn(__fsqrt_rn(__fmul_rn(__fmul_rn(2.0f, sw.rad_eff), hs)),
                                      __fsqrt_rn(__fmul_rn(__fmul_rn(2.0f, sw.rad_eff), ht)));
        if (d_2d > reach) return AW_BEYOND_HORIZON;
    }
    {                                                   // target_is_outside_vertical, :583-593
        const float rel_brg = __fsub_rn(azi, h.dir);
        const float geom_elev = aw_atan2f(dz, d_2d);
        const float apparent = __fsub_rn(geom_elev, __fmul_rn(h.rol, aw_sinf(rel_brg)));
        if ((apparent < sw.lo) || (apparent > sw.hi)) return AW_NADIR_HOLE;
    }
    return 16u;
}

// target_attempt_detection up to the draw, tut_5_1.c:643-686
__device__ __forceinline__ float aw_detection_probability(float sa, float ref_range, float ref_rcs, float ta, float tcx, float d_3d)
{
    const float r = fmaxf(1.0f, d_3d);
    const float snr = __fmul_rn(aw_powf(__fdiv_rn(ref_range, r), 4.0f), __fdiv_rn(tcx, ref_rcs));
    const float bv = ta < 400.0f ? 0.2f : (ta < 1000.0f ? 0.3f : 0.9f);
    const float dz = __fsub_rn(sa, ta);
    float sin_grazing = 0.0f;
    if (dz > 0.0f) sin_grazing = fminf(1.0f, __fdiv_rn(dz, r));
    const float clutter = __fsub_rn(1.0f, __fmul_rn(sin_grazing, 0.8f));
    const float sinr = __fmul_rn(__fmul_rn(snr, bv), clutter);
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, aw_expf(__fmul_rn(-0.5f, __fsub_rn(sinr, 10.0f)))));
}

template <bool TRACE>
__global__ void __launch_bounds__(AWACS_BLOCK)
awacs_kernel(const AwacsArgs a)
{
    __shared__ ZigHot hot;
    __shared__ uint16_t list_smem[AWACS_BLOCK / 32][AWACS_STRIDE];
    __shared__ __align__(128) float tile_smem[AWACS_BLOCK / 32][2][AWACS_STREAM_COLS][AWACS_TILE_TARGETS];
    __shared__ __align__(8) uint64_t tile_bar[AWACS_BLOCK / 32][2];

    stage_zig_hot(hot, false);
    if ((threadIdx.x & 31u) == 0u) {
        tma::barrier_init(&tile_bar[threadIdx.x >> 5][0], 1u);
        tma::barrier_init(&tile_bar[threadIdx.x >> 5][1], 1u);
        tma::barrier_init_fence();
    }
    __syncthreads();

    constexpr unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned warp = threadIdx.x >> 5;
    const uint64_t trial = (uint64_t)blockIdx.x * (AWACS_BLOCK / 32) + warp;
    if (trial >= a.num_trials) {
        return;
    }
    uint16_t *const list = list_smem[warp];
    uint32_t tile_phase = 0u;                           // bit b = parity of the next completion of tile_bar[warp][b]
    AwacsState S(a.state + trial * AWACS_STATE_BYTES);
    const AwacsTerrain ter = a.ter;
    const double PI = 3.14159265358979323846;
    const float TWO_PI_F = __fmul_rn(2.0f, (float)PI);

    Sfc64 rng;
    rng.seed(fmix64(a.master_seed, a.first_trial + trial));

    // ---- run_trial, tut_5_1.c:1139-1225: 1000 x cmb_process_start, the radar, the end event, the progress bar
    for (unsigned i = lane; i < (unsigned)AWACS_STRIDE; i += 32u) {
        S.x[i] = 0.0f; S.y[i] = 0.0f; S.alt[i] = 0.0f; S.dir[i] = 0.0f; S.vel[i] = 0.0f; S.time_s[i] = 0.0f; S.rcs_now[i] = 0.0f;
        S.flags[i] = 0u;
        S.wake_key[i] = i < (unsigned)AWACS_TARGETS ? i + 1u : 0u;
        S.wake_t[i] = i < (unsigned)AWACS_TARGETS ? 0.0 : __longlong_as_double(0x7ff0000000000000LL);
    }
    __syncwarp();
    uint32_t issued = (uint32_t)AWACS_TARGETS;
    double radar_t = 0.0;
    uint32_t radar_key = ++issued;
    bool radar_started = false;
    double end_t = a.t_end_s;
    uint32_t end_key = ++issued;
    double bar_t = 0.0;
    uint32_t bar_key = ++issued;
    uint32_t bar_cycles = 0u;
    double bar_incr = 0.0;
    bool targets_live = true;

    AwacsPlatform host = { 0.0f, 0.0f, 0.0f, 0.0f, 0.0f };
    // sensor_initialize, :1045-1065 with :1200-1206
    const float max_elev = (float)(60.0 * (2.0 * PI / 360.0));
    const float min_elev = (float)(-20.0 * (2.0 * PI / 360.0));
    const float ref_range_m = __fmul_rn(150.0f, (float)1852.0);
    const float ref_rcs = 1.0f;
    float cur_dir = (float)(PI / 2.0);
    const float rot_inc = (float)((double)__fmul_rn(6.0f, __fdiv_rn(1.0f, 60.0f)) * ((double)2.0f * PI));

    // earliest target event: (time bits, key, index); times are >= 0 so the bit patterns order like the values
    unsigned long long tmin_t = 0ull;
    uint32_t tmin_key = 1u, tmin_idx = 0u;
    bool tmin_stale = true;

    double now = 0.0;
    uint64_t pops = 0u;
    uint32_t status = TRIAL_OK;
    uint64_t lookups = 0u;                              // terrain cells this lane read while ray-marching
    const uint64_t pop_limit = (uint64_t)fmin(fmax(a.t_end_s, 0.0), 1.0e9) * 8u + 200000u;

    for (;;) {
        if (targets_live && tmin_stale) {               // 32-lane scan of the targets' pending events
            unsigned long long bt = ~0ull;
            uint32_t bk = 0xffffffffu, bi = 0u;
            for (unsigned i = lane; i < (unsigned)AWACS_TARGETS; i += 32u) {
                const unsigned long long t = (unsigned long long)__double_as_longlong(S.wake_t[i]);
                const uint32_t k = S.wake_key[i];
                if (goes_before(t, k, bt, bk)) { bt = t; bk = k; bi = i; }
            }
            const Picked p = warp_first(bt, bk);
            tmin_t = p.t;
            tmin_key = p.key;
            tmin_idx = __shfl_sync(FULL, bi, p.lane);
            tmin_stale = false;
        }
        // ---- pop-min over the four owners (time asc, key asc; every priority is 0)
        const unsigned long long NONE = ~0ull;
        unsigned long long bt = targets_live ? tmin_t : NONE;
        uint32_t bk = tmin_key;
        int kind = 0;
        if (radar_key != 0u) {
            const unsigned long long t = (unsigned long long)__double_as_longlong(radar_t);
            if (bt == NONE || goes_before(t, radar_key, bt, bk)) { bt = t; bk = radar_key; kind = 1; }
        }
        if (end_key != 0u) {
            const unsigned long long t = (unsigned long long)__double_as_longlong(end_t);
            if (bt == NONE || goes_before(t, end_key, bt, bk)) { bt = t; bk = end_key; kind = 2; }
        }
        if (bar_key != 0u) {
            const unsigned long long t = (unsigned long long)__double_as_longlong(bar_t);
            if (bt == NONE || goes_before(t, bar_key, bt, bk)) { bt = t; bk = bar_key; kind = 3; }
        }
        if (bt == NONE) {
            break;                                      // the list ran dry: cmb_event_queue_execute returns
        }
        if (pops >= pop_limit) {                        // cannot happen in a sane run; never spin on a broken one
            status |= TRIAL_ERR_KEY_OVERFLOW;
            break;
        }
        now = __longlong_as_double((long long)bt);
        if (TRACE) {
            if (lane == 0u && pops < a.trace_cap) {
                a.trace_key[trial * a.trace_cap + pops] = bk;
                a.trace_time[trial * a.trace_cap + pops] = now;
            }
        }
        pops++;

        if (kind == 0) {
            // ================= a target's resume point, target_proc :385-446 (warp-uniform; lane 0 stores)
            const uint32_t i = tmin_idx;
            uint32_t fl = S.flags[i];
            uint32_t mode = fl & AW_F_MODE;
            float gx = S.x[i], gy = S.y[i], galt = S.alt[i], gdir = S.dir[i], gvel = S.vel[i], grcs = S.rcs_now[i];
            bool to_hold = false;                       // run the loop top (hiding or driving branch) next
            double wake = 0.0;
            if (!(fl & AW_F_STARTED)) {                 // first entry, :393-402
                gx = (float)rng.uniform((double)ter.x_min, (double)ter.x_max);
                gy = (float)rng.uniform((double)ter.y_min, (double)ter.y_max);
                galt = __fadd_rn(aw_elevation(ter, gx, gy), 2.0f);
                const double ph = (double)__fdiv_rn(aw_dwell(AW_HIDING), __fadd_rn(aw_dwell(AW_HIDING), aw_dwell(AW_DRIVING)));
                mode = rng.bernoulli(ph) ? AW_HIDING : AW_DRIVING;
                fl = AW_F_STARTED;                      // tds = UNDETERMINED, not found
                to_hold = true;
            }
            else if (mode == AW_HIDING) {               // unmask, :414-420
                mode = AW_STAGING;
                grcs = aw_rcs(AW_STAGING);
                const double t_m = (double)__fdiv_rn(aw_dwell(AW_STAGING), (float)10u);
                wake = __dadd_rn(now, rng.erlang(hot, 10u, t_m));
            }
            else if (mode == AW_STAGING) {              // shoot, :422-429
                mode = AW_FIRING;
                grcs = aw_rcs(AW_FIRING);
                const double t_m = (double)__fdiv_rn(aw_dwell(AW_FIRING), (float)20u);
                wake = __dadd_rn(now, rng.erlang(hot, 20u, t_m));
            }
            else if (mode == AW_FIRING) {               // scoot, :430
                mode = AW_DRIVING;
                to_hold = true;
            }
            else {                                      // done driving, :443
                mode = AW_HIDING;
                to_hold = true;
            }
            if (to_hold) {
                if (mode == AW_HIDING) {                // :405-411
                    grcs = aw_rcs(AW_HIDING);
                    gvel = 0.0f;
                    wake = __dadd_rn(now, rng.exponential(hot, (double)aw_dwell(AW_HIDING)));
                }
                else {                                  // :432-442
                    grcs = aw_rcs(AW_DRIVING);
                    gdir = (float)rng.uniform(0.0, 2.0 * PI);
                    gvel = (float)rng.uniform(5.0, 20.0);
                    const double t_m = (double)__fdiv_rn(aw_dwell(AW_DRIVING), (float)5u);
                    wake = __dadd_rn(now, rng.erlang(hot, 5u, t_m));
                }
            }
            issued++;
            if (lane == 0u) {
                S.x[i] = gx; S.y[i] = gy; S.alt[i] = galt; S.dir[i] = gdir; S.vel[i] = gvel; S.rcs_now[i] = grcs;
                S.time_s[i] = (float)now;
                S.flags[i] = (fl & ~AW_F_MODE) | mode;
                S.wake_t[i] = wake;
                S.wake_key[i] = issued;
            }
            __syncwarp();
            tmin_stale = true;
        }
        else if (kind == 1) {
            // ================= the radar, sensor_proc :944-1036
            if (!radar_started) {
                radar_started = true;
                aw_platform_at(host, a.orbit, now);
                radar_t = __dadd_rn(now, 1.0);
                radar_key = ++issued;
                continue;
            }
            const float prev_hdg = host.dir;
            AwacsSweep sw;
            sw.prev_dir = cur_dir;
            aw_platform_at(host, a.orbit, now);
            const float ddir = __fsub_rn(host.dir, prev_hdg);
            float sweep_width = __fadd_rn(rot_inc, ddir);
            cur_dir = __fadd_rn(cur_dir, sweep_width);
            while (cur_dir >= TWO_PI_F) cur_dir = __fsub_rn(cur_dir, TWO_PI_F);
            while (cur_dir < 0.0f) cur_dir = __fadd_rn(cur_dir, TWO_PI_F);
            if (sweep_width < 0.0f) sweep_width = 0.01f;
            sw.width = sweep_width;
            sw.rad_eff = a.orbit.rad_eff;
            sw.lo = min_elev;
            sw.hi = max_elev;

            // ---- pass A: dead reckoning + the cheap tests, 32 targets at a time; survivors listed in target order.
            // The six state columns it reads arrive through shared memory: lane 0 asks the TMA engine for tile t + 2
            // (six 512-byte runs, one mbarrier) as soon as the warp is done with tile t, so the L2 / HBM latency of the
            // state block is hidden behind two tiles of geometry and costs no registers.  What this pass writes back
            // (position and time of the moving targets, verdicts) goes to the state block with ordinary stores; the
            // proxy fence below orders last tick's stores before this tick's bulk reads.
            uint32_t n_surv = 0u;
            __syncwarp();
            if (lane == 0u) {
                tma::fence_global_to_async();
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    tma::barrier_expect(&tile_bar[warp][t], (uint32_t)(AWACS_STREAM_COLS * AWACS_TILE_TARGETS * sizeof(float)));
#pragma unroll
                    for (int c = 0; c < AWACS_STREAM_COLS; c++) {
                        tma::load(tile_smem[warp][t][c], S.x + c * AWACS_STRIDE + t * AWACS_TILE_TARGETS,
                                  (uint32_t)(AWACS_TILE_TARGETS * sizeof(float)), &tile_bar[warp][t]);
                    }
                }
            }
            for (int t = 0; t < AWACS_TILES; t++) {
                const int buf = t & 1;
                tma::barrier_wait(&tile_bar[warp][buf], (tile_phase >> buf) & 1u);
                tile_phase ^= 1u << buf;
                const float (*col)[AWACS_TILE_TARGETS] = tile_smem[warp][buf];
                for (unsigned sub = 0u; sub < (unsigned)AWACS_TILE_TARGETS; sub += 32u) {
                    const unsigned j = sub + lane;
                    const unsigned i = (unsigned)t * AWACS_TILE_TARGETS + j;
                    bool survivor = false;
                    if (i < (unsigned)AWACS_TARGETS) {
                        float tx = col[0][j], ty = col[1][j], ta = col[2][j];
                        const float vel = col[4][j];
                        if (vel > 0.0f) {               // target_position_update, :507-545
                            const float dir = col[3][j];
                            const double dt = __dsub_rn(now, (double)col[5][j]);
                            const double run = __dmul_rn(dt, (double)vel);
                            float x = __fadd_rn(tx, (float)__dmul_rn(run, (double)aw_cosf(dir)));
                            if (x > ter.x_max) x = __fadd_rn(ter.x_min, __fsub_rn(x, ter.x_max));
                            else if (x < ter.x_min) x = __fsub_rn(ter.x_max, __fsub_rn(ter.x_min, x));
                            float y = __fadd_rn(ty, (float)__dmul_rn(run, (double)aw_sinf(dir)));
                            if (y > ter.y_max) y = __fadd_rn(ter.y_min, __fsub_rn(y, ter.y_max));
                            else if (y < ter.y_min) y = __fsub_rn(ter.y_max, __fsub_rn(ter.y_min, y));
                            ta = __fadd_rn(aw_elevation(ter, x, y), 2.0f);
                            tx = x;
                            ty = y;
                            S.time_s[i] = (float)now;
                            S.x[i] = tx;
                            S.y[i] = ty;
                            S.alt[i] = ta;
                        }
                        const uint32_t verdict = aw_cheap_tests(host, sw, tx, ty, ta);
                        if (verdict == 16u) {
                            survivor = true;
                        }
                        else if (verdict != 0u) {
                            S.flags[i] = (S.flags[i] & ~AW_F_TDS) | (verdict << AW_F_TDS_SHIFT);
                        }
                    }
                    const unsigned m = __ballot_sync(FULL, survivor);
                    if (survivor) {
                        list[n_surv + __popc(m & ((1u << lane) - 1u))] = (uint16_t)i;
                    }
                    n_surv += __popc(m);
                }
                __syncwarp();                           // every lane is done reading this buffer
                if (lane == 0u && t + 2 < AWACS_TILES) {
                    tma::barrier_expect(&tile_bar[warp][buf], (uint32_t)(AWACS_STREAM_COLS * AWACS_TILE_TARGETS * sizeof(float)));
#pragma unroll
                    for (int c = 0; c < AWACS_STREAM_COLS; c++) {
                        tma::load(tile_smem[warp][buf][c], S.x + c * AWACS_STRIDE + (t + 2) * AWACS_TILE_TARGETS,
                                  (uint32_t)(AWACS_TILE_TARGETS * sizeof(float)), &tile_bar[warp][buf]);
                    }
                }
            }
            __syncwarp();

            // ---- pass B: target_is_terrain_shielded (:598-638), one line of sight marched by the whole warp.
            // The reference walks the ray in half-cell steps k = 1 .. steps-1 and returns at the first step whose ray
            // altitude is below the cell under it, i.e. "shielded" = "some step is".  A ray from 9.4 km down to a target
            // spends most of its length far above any ground, so the steps are taken in stretches of
            // 32 x AWACS_CHUNK (128): first every lane brackets ONE stretch - the lowest ray altitude of its steps
            // against the highest cell they can be over - and only stretches that fail that test are marched.  The
            // bracket is exact, not approximate: step k's altitude host.alt + dz * (k * inv) and its clamped x and y are
            // each a chain of correctly rounded operations that are monotone in k, so over steps k0..k1 the altitude is
            // never below min(alt(k0), alt(k1)) and column and row stay between those of k0 and k1; a stretch is at
            // most 128 cells long, so it lies in at most 2 x 2 tiles of the 128 x 128 tile-maximum map (more: march it).
            // If that minimum is not below those tiles' maximum, no step of the stretch can be below its cell.
            uint32_t n_cand = 0u;
            for (uint32_t s = 0u; s < n_surv; s++) {
                const unsigned i = list[s];
                const float tx = S.x[i], ty = S.y[i], ta = S.alt[i];
                const float dx = __fsub_rn(tx, host.x), dy = __fsub_rn(ty, host.y), dz = __fsub_rn(ta, host.alt);
                const float d_2d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
                const float step = __fmul_rn(fminf(ter.x_scale, ter.y_scale), 0.5f);
                const int steps = (int)__fdiv_rn(d_2d, step);
                bool shielded = false;
                if (steps >= 1) {
                    const float inv = __fdiv_rn(1.0f, (float)steps);
                    constexpr int STRETCH = 32 * AWACS_CHUNK;
                    for (int sbase = 1; sbase < steps && !shielded; sbase += 32 * STRETCH) {
                        bool touch = false;
                        {
                            const int k0 = sbase + (int)lane * STRETCH;
                            if (k0 < steps) {
                                const int k1 = (k0 + STRETCH - 1 < steps - 1) ? k0 + STRETCH - 1 : steps - 1;
                                const float f0 = __fmul_rn((float)k0, inv), f1 = __fmul_rn((float)k1, inv);
                                const float low = fminf(__fadd_rn(host.alt, __fmul_rn(dz, f0)), __fadd_rn(host.alt, __fmul_rn(dz, f1)));
                                uint32_t c0, r0, c1, r1;
                                aw_col_row(ter, fmaxf(ter.x_min, fminf(__fadd_rn(host.x, __fmul_rn(dx, f0)), ter.x_max)),
                                           fmaxf(ter.y_min, fminf(__fadd_rn(host.y, __fmul_rn(dy, f0)), ter.y_max)), c0, r0);
                                aw_col_row(ter, fmaxf(ter.x_min, fminf(__fadd_rn(host.x, __fmul_rn(dx, f1)), ter.x_max)),
                                           fmaxf(ter.y_min, fminf(__fadd_rn(host.y, __fmul_rn(dy, f1)), ter.y_max)), c1, r1);
                                const uint32_t tc0 = (c0 < c1 ? c0 : c1) >> AWACS_TILE_SHIFT, tc1 = (c0 < c1 ? c1 : c0) >> AWACS_TILE_SHIFT;
                                const uint32_t tr0 = (r0 < r1 ? r0 : r1) >> AWACS_TILE_SHIFT, tr1 = (r0 < r1 ? r1 : r0) >> AWACS_TILE_SHIFT;
                                if (tc1 - tc0 > 1u || tr1 - tr0 > 1u) {
                                    touch = true;
                                }
                                else {
                                    const float *tm = ter.tile_max;
                                    const float top = fmaxf(fmaxf(__ldg(tm + (size_t)tr0 * ter.tcols + tc0), __ldg(tm + (size_t)tr0 * ter.tcols + tc1)),
                                                            fmaxf(__ldg(tm + (size_t)tr1 * ter.tcols + tc0), __ldg(tm + (size_t)tr1 * ter.tcols + tc1)));
                                    touch = low < top;
                                }
                            }
                        }
#ifdef AWACS_MARCH_EVERYTHING                                  // A/B: the round-1 march, every stretch walked
                        touch = sbase + (int)lane * STRETCH < steps;
#endif
                        unsigned todo = __ballot_sync(FULL, touch);
                        // a stretch = 32 x AWACS_CHUNK consecutive steps; each lane takes AWACS_CHUNK CONSECUTIVE ones, so the
                        // two or three half-cell steps that fall into one map cell cost one read, and a lane's successive
                        // cells share 32-byte sectors in L1 instead of every step asking L2 for its own sector
                        while (todo != 0u && !shielded) {
                            const int first = sbase + (__ffs(todo) - 1) * STRETCH;
                            todo &= todo - 1u;
                            uint32_t cell[AWACS_CHUNK];
                            float ray_alt[AWACS_CHUNK], ground[AWACS_CHUNK];
                            bool live[AWACS_CHUNK];
#pragma unroll
                            for (int u = 0; u < AWACS_CHUNK; u++) {
                                const int k = first + (int)lane * AWACS_CHUNK + u;
                                live[u] = k < steps;
                                const float f = __fmul_rn((float)k, inv);
                                float cx = __fadd_rn(host.x, __fmul_rn(dx, f));
                                float cy = __fadd_rn(host.y, __fmul_rn(dy, f));
                                ray_alt[u] = __fadd_rn(host.alt, __fmul_rn(dz, f));
                                cx = fmaxf(ter.x_min, fminf(cx, ter.x_max));
                                cy = fmaxf(ter.y_min, fminf(cy, ter.y_max));
                                cell[u] = aw_cell(ter, cx, cy);
                            }
#pragma unroll
                            for (int u = 0; u < AWACS_CHUNK; u++) {
                                const bool fresh = live[u] && (u == 0 || cell[u] != cell[u - 1]);
                                ground[u] = fresh ? __ldg(ter.map + cell[u]) : 0.0f;
                                lookups += fresh ? 1u : 0u;
                            }
                            bool hit = false;
#pragma unroll
                            for (int u = 0; u < AWACS_CHUNK; u++) {
                                if (u > 0 && cell[u] == cell[u - 1]) ground[u] = ground[u - 1];
                                hit |= live[u] && (ray_alt[u] < ground[u]);
                            }
                            shielded = __any_sync(FULL, hit);
                        }
                    }
                }
                if (shielded) {
                    if (lane == 0u) S.flags[i] = (S.flags[i] & ~AW_F_TDS) | (AW_TERRAIN_SHIELDED << AW_F_TDS_SHIFT);
                }
                else {
                    if (lane == 0u) list[n_cand] = (uint16_t)i;     // n_cand <= s: never overtakes the reader
                    n_cand++;
                }
                __syncwarp();
            }

            // ---- pass C: target_attempt_detection (:643-689); probabilities in parallel, draws in target order
            for (uint32_t base = 0u; base < n_cand; base += 32u) {
                const uint32_t mine = base + lane;
                unsigned i = 0u;
                float pd = 0.0f;
                uint32_t fl = 0u;
                if (mine < n_cand) {
                    i = list[mine];
                    const float tx = S.x[i], ty = S.y[i], ta = S.alt[i];
                    const float dx = __fsub_rn(tx, host.x), dy = __fsub_rn(ty, host.y), dz = __fsub_rn(ta, host.alt);
                    const float d_2d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
                    const float d_3d = __fsqrt_rn(__fadd_rn(__fmul_rn(d_2d, d_2d), __fmul_rn(dz, dz)));
                    pd = aw_detection_probability(host.alt, ref_range_m, ref_rcs, ta, S.rcs_now[i], d_3d);
                    fl = S.flags[i];
                }
                const uint32_t count = min(32u, n_cand - base);
                bool seen = false;
                for (uint32_t k = 0u; k < count; k++) {
                    const double u = rng.uniform01();   // cmb_random_bernoulli(pd): cmb_random() <= p
                    if (lane == k) seen = u <= (double)pd;
                }
                if (mine < n_cand) {
                    fl = (fl & ~AW_F_TDS) | ((seen ? AW_DETECTED : AW_MISSED) << AW_F_TDS_SHIFT);
                    if (seen && (fl & AW_F_MODE) != AW_FIRING) fl |= AW_F_FOUND;
                    S.flags[i] = fl;
                }
            }
            __syncwarp();
            radar_t = __dadd_rn(now, 1.0);
            radar_key = ++issued;
        }
        else if (kind == 2) {
            // ================= end_sim, :1100-1112: cmb_process_stop on the radar and on every target
            end_key = 0u;
            radar_key = 0u;
            targets_live = false;
        }
        else {
            // ================= ent_proc, :1118-1137: 100 holds of a hundredth of the run
            if (bar_cycles == 0u) {
                bar_incr = __ddiv_rn(__dsub_rn(a.t_end_s, now), 100.0);
            }
            if (bar_cycles < 100u) {
                bar_cycles++;
                bar_t = __dadd_rn(now, bar_incr);
                bar_key = ++issued;
            }
            else {
                bar_key = 0u;
            }
        }
    }

    // ---- results: struct trial.num_found (:1227-1232) and the final picture of the targets
    uint32_t found = 0u, tds_n[6] = { 0u, 0u, 0u, 0u, 0u, 0u }, mode_n[4] = { 0u, 0u, 0u, 0u };
    for (unsigned i = lane; i < (unsigned)AWACS_TARGETS; i += 32u) {
        const uint32_t fl = S.flags[i];
        found += (fl & AW_F_FOUND) ? 1u : 0u;
        const uint32_t tds = (fl & AW_F_TDS) >> AW_F_TDS_SHIFT;
#pragma unroll
        for (uint32_t k = 0u; k < 6u; k++) tds_n[k] += tds == k ? 1u : 0u;
#pragma unroll
        for (uint32_t k = 0u; k < 4u; k++) mode_n[k] += (fl & AW_F_MODE) == k ? 1u : 0u;
    }
    found = __reduce_add_sync(FULL, found);
    uint64_t marched = lookups;
    for (int o = 16; o > 0; o >>= 1) marched += __shfl_xor_sync(FULL, marched, o);
#pragma unroll
    for (uint32_t k = 0u; k < 6u; k++) tds_n[k] = __reduce_add_sync(FULL, tds_n[k]);
#pragma unroll
    for (uint32_t k = 0u; k < 4u; k++) mode_n[k] = __reduce_add_sync(FULL, mode_n[k]);
    if (lane == 0u) {
        double sum_x = 0.0, sum_y = 0.0;                // in target order, as the oracle adds them
        for (unsigned i = 0u; i < (unsigned)AWACS_TARGETS; i++) {
            sum_x = __dadd_rn(sum_x, (double)S.x[i]);
            sum_y = __dadd_rn(sum_y, (double)S.y[i]);
        }
        if (a.events)    a.events[trial] = pops;
        if (a.objects)   a.objects[trial] = found;
        if (a.t_end)     a.t_end[trial] = now;
        if (a.sum_wait)  a.sum_wait[trial] = sum_x;
        if (a.status)    a.status[trial] = status;
        if (a.max_queue) a.max_queue[trial] = (uint32_t)AWACS_TARGETS + 3u;
        if (a.counters) {
            uint64_t *c = a.counters + trial * 8u;
#pragma unroll
            for (uint32_t k = 0u; k < 6u; k++) c[k] = tds_n[k];
            c[6] = (uint64_t)mode_n[0] | ((uint64_t)mode_n[1] << 16) | ((uint64_t)mode_n[2] << 32) | ((uint64_t)mode_n[3] << 48);
            c[7] = marched;
            (void)sum_y;
        }
    }
}

#ifndef AWACS_HOST_EMULATION
// The tile-maximum map of a terrain: one CTA per 128 x 128 tile, rows read as coalesced 512-byte segments.
__global__ void __launch_bounds__(256)
aw_tile_max_kernel(const float *map, uint32_t cols, uint32_t rows, float *tile_max, uint32_t tcols)
{
    __shared__ float part[8];
    const uint32_t c0 = blockIdx.x << AWACS_TILE_SHIFT, r0 = blockIdx.y << AWACS_TILE_SHIFT;
    float top = -3.402823466e+38f;                      // a caller's map may hold negative elevations
    for (uint32_t e = threadIdx.x; e < AWACS_TILE * AWACS_TILE; e += blockDim.x) {
        const uint32_t r = r0 + (e >> AWACS_TILE_SHIFT), c = c0 + (e & (AWACS_TILE - 1u));
        if (r < rows && c < cols) top = fmaxf(top, __ldg(map + (size_t)r * cols + c));
    }
    for (int o = 16; o > 0; o >>= 1) top = fmaxf(top, __shfl_xor_sync(0xffffffffu, top, o));
    if ((threadIdx.x & 31u) == 0u) part[threadIdx.x >> 5] = top;
    __syncthreads();
    if (threadIdx.x == 0u) {
        for (int w = 1; w < 8; w++) top = fmaxf(top, part[w]);
        tile_max[(size_t)blockIdx.y * tcols + blockIdx.x] = top;
    }
}
#endif

}  // namespace cimba_b200
