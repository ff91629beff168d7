// gg1_fast.cuh - the G/G/1 replication kernel (BASELINE config 4: Erlang-2 arrivals,
// ziggurat-normal service times redrawn while negative), in the straight-line predicated
// form of mm1_fast.cuh.
//
// Same model, same per-trial arithmetic, the same order of random draws, key issues and
// pops as queue_kernel<1> in queue_model.cuh (the readable formulation, kept as
// job->variant = 1); see that file for the mapping onto the reference.  What differs is how
// a warp executes it:
//   * one event step is a single predicated instruction sequence for all 32 lanes (pop-min
//     over the two process-owned event slots, arrival body, service body, hold);
//   * the generator runs TWO raw sfc64 outputs ahead of the simulation.  An arrival's
//     cmb_random_erlang(2, m/2) consumes both (two hot exponentials), a service's
//     cmb_random_normal consumes the first; the table look-ups and the 64-bit -> double
//     conversions of the look-ahead happen a step early, off the pop -> push chain;
//   * a draw that leaves the ziggurats' rectangles (1.6 % of exponentials, 1.2 % of normals)
//     or a negative service time parks the lane.  Parked lanes are served in batches by the
//     reference-order slow path: the generator is rewound by the two outputs of look-ahead
//     (Sfc64::rewind - sfc64's transition is a bijection) and the plain cmb_random_erlang /
//     cmb_random_normal code of rng.cuh runs from there, so the stream position afterwards
//     is exactly the reference's.
#pragma once

#include "engine.cuh"
#include "mm1_fast.cuh"
#include "queue_model.cuh"
#include "rng.cuh"

namespace cimba_b200 {

#ifndef GG1_PARK_MASK
#define GG1_PARK_MASK 7u
#endif
#ifndef GG1_COLD_BATCH
#define GG1_COLD_BATCH 4
#endif

template <bool TRACE>
__global__ void __launch_bounds__(QUEUE_BLOCK)
gg1_kernel(const QueueArgs a)
{
    __shared__ ZigHot hot;                              // both layer-width tables
    __shared__ double ring_smem[QUEUE_WINDOW * QUEUE_BLOCK];
    __shared__ double scratch_smem[QUEUE_BLOCK];        // sink for the store of lanes that do not put

    stage_zig_hot(hot, true);
    __syncthreads();

    constexpr unsigned FULL = 0xffffffffu;
    constexpr uint32_t WMASK = QUEUE_WINDOW - 1;
    constexpr uint32_t ROW = QUEUE_BLOCK * 8u;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const unsigned lane = threadIdx.x & 31u;
    const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t trial;
    bool alive;
    if (a.mapping == 32) {
        trial = gtid >> 5;
        alive = (lane == 0u) && (trial < a.num_trials);
    }
    else {
        trial = gtid;
        alive = trial < a.num_trials;
    }

    Sfc64 rng;
    rng.a = rng.b = rng.c = rng.d = 0u;
    double t_arr = INF, t_srv = INF;                    // the two process-owned event slots (mm1_fast.cuh)
    uint32_t k_arr = 0u, k_srv = 0u;
    uint32_t issued = 0u;
    double now = 0.0, stamp = 0.0, sum_wait = 0.0;
    double arr_half = 0.5, srv_mean = 1.0, srv_sigma = 0.25;
    uint32_t produced = 0u, served = 0u, dropped = 0u, status = TRIAL_OK, longest = 0u;
    const uint32_t quota = (uint32_t)a.num_objects;
    uint32_t win = (uint32_t)__cvta_generic_to_shared(&ring_smem[threadIdx.x]);
    uint32_t tab_e = (uint32_t)__cvta_generic_to_shared(&hot.exp_x[0]);
    uint32_t tab_n = (uint32_t)__cvta_generic_to_shared(&hot.nor_x[0]);
    uint32_t scratch = (uint32_t)__cvta_generic_to_shared(&scratch_smem[threadIdx.x]);
    asm volatile("" : "+r"(win), "+r"(tab_e), "+r"(tab_n), "+r"(scratch));
    double *const spill = (a.spill_cap && alive) ? a.spill + trial * a.spill_cap : nullptr;
    const uint32_t spill_mask = a.spill_cap - 1u;

    // two raw outputs of look-ahead, with their hot-path variates already formed:
    // e = as a standard exponential, z = as a standard normal (the sign rides in the integer)
    uint64_t u1 = 0u, u2 = 0u;
    double e1 = 0.0, z1 = 0.0, e2 = 0.0, z2 = 0.0;
    uint32_t pops = 0u;

#define GG1_FORM(u, e, z)                                                                           \
    do {                                                                                            \
        const uint32_t idx_ = ((uint32_t)(u) & 0xffu) * 8u;                                         \
        (e) = __dmul_rn(lds_f64(tab_e + idx_), __ull2double_rn(u));                                 \
        (z) = __dmul_rn(lds_f64(tab_n + idx_), __ll2double_rn((long long)(u)));                     \
    } while (0)

    if (alive) {
        arr_half = __dmul_rn(0.5, a.arr_mean[trial]);   // cmb_random_erlang(2, 0.5 * mean)
        srv_mean = a.srv_mean[trial];
        srv_sigma = __dmul_rn(0.25, srv_mean);          // cmb_random_normal(mean, 0.25 * mean)
        rng.seed(fmix64(a.master_seed, a.first_trial + trial));
        t_arr = 0.0; k_arr = pack_key(1u, ACT_START);
        t_srv = 0.0; k_srv = pack_key(2u, ACT_START);
        issued = 2u;
        u1 = rng.next();
        u2 = rng.next();
        GG1_FORM(u1, e1, z1);
        GG1_FORM(u2, e2, z2);
    }

    // bit 0 alive, bit 1 parked, bit 2 the parked draw belongs to the arrival process
    uint32_t flags = alive ? 1u : 0u;
    uint32_t step = 0u;

    while (__any_sync(FULL, flags & 1u)) {
        // ---------------- pop-min
        const bool go0 = (flags & 3u) == 1u;
        const bool first_arr = (t_arr < t_srv) | ((t_arr == t_srv) & (k_arr < k_srv));
        const uint32_t key = first_arr ? k_arr : k_srv;
        const uint32_t act = key & 3u;
        const bool go = go0 & (key != 0u);
        const bool done = go0 & (key == 0u);
        const bool is_arr = go & first_arr;
        const bool is_srv = go & !first_arr;
        const bool wake = act == ACT_WAKE_TIME;
        if (TRACE) {
            if (go && pops < a.trace_cap) {
                a.trace_key[trial * a.trace_cap + pops] = key >> 2;
                a.trace_time[trial * a.trace_cap + pops] = first_arr ? t_arr : t_srv;
            }
            pops += go ? 1u : 0u;
        }
        if (go) now = first_arr ? t_arr : t_srv;

        // ---------------- arrival body: back from hold -> put
        const uint32_t q_len = produced - served;
        const bool put = is_arr & wake;
        const bool put_far = put & (q_len >= (uint32_t)QUEUE_WINDOW);
        sts_f64((put & !put_far) ? win + (produced & WMASK) * ROW : scratch, now);
        if (put_far) {
            if (spill != nullptr && q_len - QUEUE_WINDOW <= spill_mask) {
                spill[produced & spill_mask] = now;
            }
            else {
                status |= TRIAL_ERR_QUEUE_OVERFLOW;
                dropped++;
                served++;
            }
        }
        if (put) produced++;
        longest = max(longest, produced - served);
        const bool ring_bell = put & (k_srv == 0u);
        if (ring_bell) {
            issued++;
            t_srv = now;
            k_srv = pack_key(issued, ACT_WAKE_RESOURCE);
        }

        // ---------------- service body
        const bool finished = is_srv & wake;
        const double new_sum = __dadd_rn(sum_wait, __dsub_rn(now, stamp));
        if (finished) sum_wait = new_sum;
        const bool take = is_srv & (produced != served);
        const uint32_t head_slot = win + (served & WMASK) * ROW;
        const double head_stamp = lds_f64(head_slot);
        if (take) stamp = head_stamp;
        if (take & (produced - served > (uint32_t)QUEUE_WINDOW)) {
            sts_f64(head_slot, spill[(served + QUEUE_WINDOW) & spill_mask]);
        }
        if (take) served++;

        // ---------------- hold: consume the look-ahead, insert the wake-up
        const bool draw_arr = is_arr & (produced < quota);
        const bool draw = take | draw_arr;
        const uint32_t i1 = (uint32_t)u1 & 0xffu, i2 = (uint32_t)u2 & 0xffu;
        // erlang-2: x = 0.0; x += m * e1; x += m * e2 (0.0 + v == v exactly)
        const double d_arr = __dadd_rn(__dmul_rn(arr_half, e1), __dmul_rn(arr_half, e2));
        const double d_srv = __dadd_rn(srv_mean, __dmul_rn(srv_sigma, z1));
        const bool hot_arr = (i1 <= ZIG_EXP_MAX) & (i2 <= ZIG_EXP_MAX);
        const bool hot_srv = (i1 <= ZIG_NOR_MAX) & !(d_srv < 0.0);
        const bool push = draw & (is_arr ? hot_arr : hot_srv);
        const double when = __dadd_rn(now, is_arr ? d_arr : d_srv);
        if (push) issued++;
        const double t_new = push ? when : INF;
        const uint32_t k_new = push ? pack_key(issued, ACT_WAKE_TIME) : 0u;
        if (is_arr) { t_arr = t_new; k_arr = k_new; }
        if (is_srv) { t_srv = t_new; k_srv = k_new; }
        if (draw & !push) flags = is_arr ? 7u : 3u;
        // refill: a service consumed one raw output, an arrival two
        if (push) {
            if (!is_arr) {
                u1 = u2; e1 = e2; z1 = z2;
            }
            else {
                u1 = rng.next();
                GG1_FORM(u1, e1, z1);
            }
            u2 = rng.next();
            GG1_FORM(u2, e2, z2);
        }

        // ---------------- rare paths
        if (done) {
            flags = 0u;
            if (a.events)    a.events[trial] = issued;
            if (a.objects)   a.objects[trial] = served - dropped;
            if (a.t_end)     a.t_end[trial] = now;
            if (a.sum_wait)  a.sum_wait[trial] = sum_wait;
            if (a.status)    a.status[trial] = status | (issued > 0x3ffffff0u ? TRIAL_ERR_KEY_OVERFLOW : 0u);
            if (a.max_queue) a.max_queue[trial] = longest;
        }
        if ((++step & GG1_PARK_MASK) != 0u) {
            continue;
        }
        const unsigned pm = __ballot_sync(FULL, flags & 2u);
        if (pm != 0u) {
            const unsigned am = __ballot_sync(FULL, flags & 1u);
            if (__popc(pm) >= GG1_COLD_BATCH || pm == am) {
                if (flags & 2u) {
                    // the reference-order slow path: hand the two outputs of look-ahead back
                    rng.rewind();
                    rng.rewind();
                    const bool parked_is_arr = (flags & 4u) != 0u;
                    double dur;
                    if (parked_is_arr) {
                        dur = rng.erlang(hot, 2u, arr_half);
                    }
                    else {
                        do {
                            dur = rng.normal(hot, srv_mean, srv_sigma);
                        } while (dur < 0.0);
                    }
                    const double at = __dadd_rn(now, dur);
                    issued++;
                    if (parked_is_arr) { t_arr = at; k_arr = pack_key(issued, ACT_WAKE_TIME); }
                    else               { t_srv = at; k_srv = pack_key(issued, ACT_WAKE_TIME); }
                    flags = 1u;
                    u1 = rng.next();
                    u2 = rng.next();
                    GG1_FORM(u1, e1, z1);
                    GG1_FORM(u2, e2, z2);
                }
            }
        }
    }
#undef GG1_FORM
    if (a.diag != nullptr && lane == 0u) {             // bench.py: loop iterations -> issued warp-instructions
        atomicAdd(a.diag, (unsigned long long)step);
        atomicAdd(a.diag + 1, 1ull);
    }
}

}  // namespace cimba_b200
