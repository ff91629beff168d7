// pool_fast.cuh - M/M/c through a cmb_resourcepool (BASELINE config 3) in predicated form.
//
// Same model, same arithmetic, the same order of random draws and key issues as
// pool_kernel in pool_model.cuh (the readable formulation, kept as job->variant = 1; see
// that file for the mapping onto the reference's acquire / release / guard mechanics).
// Profiling pool_kernel at config 3's 32 768 trials per GPU showed ~500 warp instructions
// per event step with 19 of 32 lanes active: the generator body, the "service finished"
// body and the "try to acquire" body are separate divergent regions, each with its own
// inlined copies of the event-list insert and the wait-list take.
//
// Here one event step is one instruction sequence for all lanes.  Whatever the event,
// at most ONE event is scheduled at the current time (an arrival starts its customer; a
// release, or a grab that leaves units over, wakes the head waiter - never both) and at
// most ONE in the future (the generator's or the customer's hold), in that order, which is
// the order the reference issues their keys in.  The wait list (arrival stamps, FIFO) is
// indexed by two counters like mm1_fast.cuh's queue: a blocked attempt stores at the tail,
// a signal reads the head; lanes that do neither store to a scratch word.
#pragma once

#include "engine.cuh"
#include "mm1_fast.cuh"
#include "pool_model.cuh"
#include "rng.cuh"

namespace cimba_b200 {

#ifndef POOL_PARK_MASK
#define POOL_PARK_MASK 3u
#endif

template <bool TRACE>
__global__ void __launch_bounds__(POOL_BLOCK)
pool_fast_kernel(const PoolArgs a)
{
    __shared__ double exp_x[256];
    __shared__ EventHead fel_head[POOL_FEL_CAP * POOL_BLOCK];
    __shared__ double fel_pay[POOL_FEL_CAP * POOL_BLOCK];
    __shared__ double wait_smem[POOL_WINDOW * POOL_BLOCK];
    __shared__ double scratch_smem[POOL_BLOCK];

    for (unsigned i = threadIdx.x; i < 256u; i += blockDim.x) {
        exp_x[i] = zig::zig_exp_x[i];
    }
    __syncthreads();

    constexpr unsigned FULL = 0xffffffffu;
    constexpr uint32_t WMASK = POOL_WINDOW - 1;
    constexpr uint32_t ROW = POOL_BLOCK * 8u;
    const uint64_t trial = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool exists = trial < a.num_trials;

    Sfc64 rng;
    rng.a = rng.b = rng.c = rng.d = 0u;
    EventList<POOL_FEL_CAP> fel;
    fel.init(&fel_head[threadIdx.x], &fel_pay[threadIdx.x], POOL_BLOCK);
    uint32_t win = (uint32_t)__cvta_generic_to_shared(&wait_smem[threadIdx.x]);
    uint32_t tab = (uint32_t)__cvta_generic_to_shared(&exp_x[0]);
    uint32_t scratch = (uint32_t)__cvta_generic_to_shared(&scratch_smem[threadIdx.x]);
    asm volatile("" : "+r"(win), "+r"(tab), "+r"(scratch));
    double *const spill = (a.spill_cap && exists) ? a.spill + trial * a.spill_cap : nullptr;
    const uint32_t spill_mask = a.spill_cap - 1u;

    double now = 0.0, sum_wait = 0.0, arr_mean = 1.0, srv_mean = 1.0;
    uint32_t pops = 0u, produced = 0u, served = 0u, status = TRIAL_OK;
    uint32_t in_use = 0u, live = 0u, most_live = 0u;
    uint32_t w_in = 0u, w_out = 0u;                     // wait list: stamps put / taken (FIFO: length = w_in - w_out)
    const uint32_t capacity = (uint32_t)a.servers;
    const uint32_t quota = (uint32_t)a.num_objects;

    // one raw output of look-ahead with its hot-path variate already formed (mm1_fast.cuh): the table
    // look-up and the 64-bit -> double conversion leave the pop -> push chain; stream order unchanged
    uint64_t u_next = 0u;
    double e_next = 0.0;
    if (exists) {
        arr_mean = a.arr_mean[trial];
        srv_mean = a.srv_mean[trial];
        rng.seed(fmix64(a.master_seed, a.first_trial + trial));
        fel.schedule(ACT_START, TAG_GENERATOR, 0.0, 0.0);       // cmb_process_start(source)
        u_next = rng.next();
        e_next = __dmul_rn(lds_f64(tab + ((uint32_t)u_next & 0xffu) * 8u), __ull2double_rn(u_next));
    }

    // bit 0 alive, bit 1 parked (a draw waits for the ziggurat slow path)
    uint32_t flags = exists ? 1u : 0u;
    uint32_t parked_tag = 0u;
    double parked_pay = 0.0;
    uint32_t step = 0u;

    while (__any_sync(FULL, flags & 1u)) {
        const bool go0 = (flags & 3u) == 1u;
        const uint32_t scan = __reduce_max_sync(FULL, go0 ? fel.count : 0u);
        EventHead ev;
        double pay = 0.0;
        ev.time = 0.0;
        ev.keyact = 0u;
        ev.tag = 0u;
        bool go = false;
        if (go0) {
            go = fel.pop(scan, ev, pay);
        }
        const bool done = go0 & !go;                    // the list ran dry: cmb_event_queue_execute returns
        if (go) now = ev.time;
        if (TRACE) {
            if (go && pops < a.trace_cap) {
                a.trace_key[trial * a.trace_cap + pops] = ev.keyact >> 2;
                a.trace_time[trial * a.trace_cap + pops] = now;
            }
        }
        pops += go ? 1u : 0u;

        // ---------------- which of the three bodies this event resumes
        const bool is_gen = ev.tag == TAG_GENERATOR;
        const bool wake = (ev.keyact & 3u) == ACT_WAKE_TIME;
        const bool arrive = go & is_gen & wake;         // generator back from hold: a customer arrives
        const bool depart = go & !is_gen & wake;        // service finished: release, account, exit
        const bool attempt = go & !is_gen & !wake;      // START or WAKE_RESOURCE: (re)try to acquire one unit

        // ---------------- the pool (src/cmb_resourcepool.c:362-533, 561-605; amounts of 1)
        const bool got = attempt & (in_use < capacity);
        const bool blocked = attempt & !got;
        in_use += (got ? 1u : 0u) - (depart ? 1u : 0u);
        // cmb_resourceguard_signal after a release, or after a grab "in case someone else can use the
        // leftovers": wake the head waiter if a unit is free
        const uint32_t w_len = w_in - w_out;
        const bool sig = (depart | got) & (w_len != 0u) & (in_use < capacity);

        // ---------------- the wait list: stamps of the customers queued at the guard
        const bool put_far = blocked & (w_len >= (uint32_t)POOL_WINDOW);
        sts_f64((blocked & !put_far) ? win + (w_in & WMASK) * ROW : scratch, pay);
        bool dropped = false;
        if (put_far) {                                  // rare: beyond the on-chip window
            if (spill != nullptr && w_len - POOL_WINDOW <= spill_mask) {
                spill[w_in & spill_mask] = pay;
            }
            else {
                status |= TRIAL_ERR_GUARD_OVERFLOW;     // entry dropped: the trial is void from here on
                dropped = true;
            }
        }
        const uint32_t head_slot = win + (w_out & WMASK) * ROW;
        const double head_stamp = lds_f64(head_slot);   // harmless when nobody waits
        if (sig & (w_len > (uint32_t)POOL_WINDOW)) {    // rare: refill the freed slot from HBM
            sts_f64(head_slot, spill[(w_out + POOL_WINDOW) & spill_mask]);
        }
        w_in += (blocked & !dropped) ? 1u : 0u;
        w_out += sig ? 1u : 0u;

        // ---------------- bookkeeping of the bodies
        if (arrive) {
            live++;
            produced++;
        }
        most_live = max(most_live, live);
        const double new_sum = __dadd_rn(sum_wait, __dsub_rn(now, pay));
        if (depart) {
            sum_wait = new_sum;
            served++;
            live--;
        }

        // ---------------- at most one event at the current time: START of the new customer, or the
        // head waiter's WAKE_RESOURCE (key issued before the hold's, as the reference does)
        const bool soon = arrive | sig;
        {
            const uint32_t k = fel.issued + (soon ? 1u : 0u);
            fel.issued = k;
            const uint32_t at = min(fel.count, (uint32_t)POOL_FEL_CAP - 1u);
            if (soon & (fel.count >= (uint32_t)POOL_FEL_CAP)) status |= TRIAL_ERR_FEL_OVERFLOW;
            if (soon) {
                EventList<POOL_FEL_CAP>::st_head(fel.head + at * fel.hstride, now,
                                                 (k << 2) | (arrive ? ACT_START : ACT_WAKE_RESOURCE), TAG_CUSTOMER);
                sts_f64(fel.pay + at * fel.pstride, arrive ? now : head_stamp);
                fel.count = min(fel.count + 1u, (uint32_t)POOL_FEL_CAP);
            }
        }

        // ---------------- the hold: generator's next inter-arrival, or the customer's service
        const bool draw = got | (go & is_gen & (produced < quota));
        {
            const uint32_t tag = is_gen ? TAG_GENERATOR : TAG_CUSTOMER;
            const double keep = is_gen ? 0.0 : pay;     // the customer's arrival stamp travels with its event
            const bool hot = Sfc64::exp_is_hot(u_next);
            const bool push = draw & hot;
            const double when = __dadd_rn(now, __dmul_rn(is_gen ? arr_mean : srv_mean, e_next));
            const uint32_t k = fel.issued + (push ? 1u : 0u);
            fel.issued = k;
            const uint32_t at = min(fel.count, (uint32_t)POOL_FEL_CAP - 1u);
            if (push & (fel.count >= (uint32_t)POOL_FEL_CAP)) status |= TRIAL_ERR_FEL_OVERFLOW;
            if (push) {
                EventList<POOL_FEL_CAP>::st_head(fel.head + at * fel.hstride, when, (k << 2) | ACT_WAKE_TIME, tag);
                sts_f64(fel.pay + at * fel.pstride, keep);
                fel.count = min(fel.count + 1u, (uint32_t)POOL_FEL_CAP);
                u_next = rng.next();                    // refill the look-ahead
                e_next = __dmul_rn(lds_f64(tab + ((uint32_t)u_next & 0xffu) * 8u), __ull2double_rn(u_next));
            }
            if (draw & !hot) {
                flags |= 2u;
                parked_tag = tag;
                parked_pay = keep;
            }
        }

        // ---------------- rare paths
        if (done) {
            flags = 0u;
            if (a.events)    a.events[trial] = pops;
            if (a.objects)   a.objects[trial] = served;
            if (a.t_end)     a.t_end[trial] = now;
            if (a.sum_wait)  a.sum_wait[trial] = sum_wait;
            if (a.status)    a.status[trial] = status | (fel.issued > 0x3ffffff0u ? TRIAL_ERR_KEY_OVERFLOW : 0u);
            if (a.max_queue) a.max_queue[trial] = most_live;
        }
        if ((++step & POOL_PARK_MASK) != 0u) {
            continue;
        }
        const unsigned pm = __ballot_sync(FULL, flags & 2u);
        if (pm != 0u) {
            const unsigned am = __ballot_sync(FULL, flags & 1u);
            if (__popc(pm) >= POOL_COLD_BATCH || pm == am) {
                if (flags & 2u) {
                    const double mean = parked_tag == TAG_GENERATOR ? arr_mean : srv_mean;
                    const double dur = __dmul_rn(mean, rng.exp_cold(u_next));
                    if (!fel.schedule(ACT_WAKE_TIME, parked_tag, __dadd_rn(now, dur), parked_pay)) {
                        status |= TRIAL_ERR_FEL_OVERFLOW;
                    }
                    flags &= ~2u;
                    u_next = rng.next();
                    e_next = __dmul_rn(lds_f64(tab + ((uint32_t)u_next & 0xffu) * 8u), __ull2double_rn(u_next));
                }
            }
        }
    }
    if (a.diag != nullptr && (threadIdx.x & 31u) == 0u) {    // bench.py: loop iterations -> issued warp-instructions
        atomicAdd(a.diag, (unsigned long long)step);
        atomicAdd(a.diag + 1, 1ull);
    }
}

}  // namespace cimba_b200
