// harbor_model.cuh - the harbor: the reference's own condition-variable test model
// (test/test_condition.c, also tutorial/tut_4_1.c), model 10.
//
// This is the "cmb_condition + cmb_resourcepool heavy, divergent-process" workload of
// BASELINE config 5 / SURVEY.md 8f-2, and the one model of this engine whose results are
// pinned by a golden file of the reference itself: with seed 0x34f05c64d7ad598f and a
// duration of 100 years (873 600 h) test/reference/condition.txt reports N 328781 small and
// 109454 large ships, mean system times 10.91 / 17.48, tug utilisation N 1736975 mean 0.8025.
//
// Processes (creation order fixes the event keys, test/test_condition.c:523-586):
//   0 weather    every hour: wind from cmb_random_rayleigh + 2 x cmb_random_PERT, then
//                cmb_condition_signal(harbormaster)
//   1 tide       every hour: depth from a sine model of cmb_time() + the wind, then signal
//   2 arrivals   hold(exp); create a ship PROCESS (25 % large) and start it
//   3 departures cmb_condition_wait(davyjones); collect the ship's exit value, free it
//   4 dots       hold(one year), for ever
//   5.. ships    wait on the harbormaster until depth, wind, tugs and a berth suit; acquire a
//                berth and 1 or 3 tugs (cmb_resourcepool, greedy partial grabs, src/
//                cmb_resourcepool.c:362-533); dock (PERT hold); release tugs; unload (PERT);
//                acquire tugs again - may queue at the pool's guard; undock; release all;
//                join the departed list; cmb_condition_signal(davyjones); return.
// and an end event that stops 0..4 and then every active ship in arrival order.
//
// cmb_condition_signal is the evaluate-all, two-pass signal of src/cmb_condition.c:120-167:
// wake-up events are scheduled in the guard's heap ARRAY order, so the guard keeps the
// reference's physical layout (BinHeap in general.cuh).  The three pools' histories
// (cmb_resourcepool_start_recording) are folded into time-weighted summaries on the fly
// (TimeWeighted, summary.cuh).  sin/fmod come from CUDA's libm; the water depth they feed
// is only ever compared against 8.0 / 13.0, so a last-place difference from glibc is
// invisible unless the depth lands within 1e-15 of a threshold.
//
// One trial per lane, state in HBM; ship structs live in a fixed table of HARBOR_SHIPS
// slots (the reference mallocs them), overflow is reported in the trial's status word.
#pragma once

#include <cmath>

#include "distributions.cuh"
#include "general.cuh"
#include "guarded_model.cuh"
#include "summary.cuh"

namespace cimba_b200 {

#ifndef HARBOR_INLINE_ON_CHIP
#define HARBOR_INLINE_ON_CHIP 1     // 0: the warp-per-trial kernel calls the out-of-line bindings too (an A/B knob: code size vs call latency)
#endif
constexpr uint32_t HARBOR_FIXED = 5u;                   // weather, tide, arrivals, departures, dots
constexpr uint16_t HARBOR_NONE = 0xffffu;
constexpr int HARBOR_BLOCK_ON_CHIP = 128;               // 4 warps = 4 trials per CTA when the state is in shared memory

// libm's sin carries a large argument-reduction slow path; four inlined copies of it (the tide model)
// are 10 % of this kernel's code.  One out-of-line copy instead.
__device__ __noinline__ double harbor_sin(double x)
{
    return sin(x);
}

struct HarborProc {
    uint8_t  pc, status, n_awaits, active;
    uint8_t  await_type[2];
    uint8_t  size, need, held_tugs, held_berth, rem, in_use;   // in_use: the table slot is allocated
    uint32_t await_ref[2];
    uint32_t hold_handle, guard_key, id;
    uint16_t next_departed;
    double   t_arr, t_sys;
};

struct HarborPool {
    uint32_t cap, in_use;
    TimeWeighted hist;
};

// Capacities are template parameters: the HBM-resident state is generous (lane per trial), the
// shared-memory-resident one (warp per trial) is sized for the tested regimes; both report an
// overflow in the trial's status word.
template <int FEL_CAP, int GUARD_CAP, int SHIPS>
struct HarborStateT {
    static constexpr int GUARD = GUARD_CAP;
    static constexpr uint32_t PROCS = HARBOR_FIXED + (uint32_t)SHIPS;
    BinHeap<FEL_CAP, EventOrder>   fel;
    BinHeap<GUARD_CAP, GuardOrder> harbormaster;
    BinHeap<GUARD_CAP, GuardOrder> tug_guard;
    BinHeap<3, GuardOrder>                davyjones;
    BinHeap<3, GuardOrder>                berth_guard[2];      // nobody ever waits here: a ship only asks when one is free
    HarborProc proc[PROCS];
    HarborPool tugs, berth[2];
    double   wind_magnitude, wind_direction, water_depth;
    uint32_t guard_seq, status, next_id, alive, most_alive;
    uint16_t departed;
};
using HarborState = HarborStateT<127, 127, 120>;        // HBM-resident, one trial per lane (~17 KB)
using HarborStateOnChip = HarborStateT<47, 47, 43>;     // shared-memory-resident, one trial per warp (~5.6 KB)

// the on-chip kernel's dynamic shared memory: [warps per CTA] HarborStateOnChip
extern __shared__ __align__(16) unsigned char harbor_smem[];

struct HarborArgs {
    int32_t  tugs;
    uint64_t master_seed, first_trial, num_trials, duration;
    const double *arr_mean, *unload_small;
    uint64_t *events, *objects;
    double   *t_end, *sum_wait;
    uint32_t *status, *max_queue;
    uint64_t *counters;
    void     *state;            // [num_trials] HarborState (lane per trial); unused by the on-chip kernel
    uint32_t  repair;           // lane-per-trial kernel: only re-run trials whose status word is non-zero
    uint64_t  trace_cap;
    uint64_t *trace_key;
    double   *trace_time;
};

// Member functions come in two bindings (the `_out` / `_impl` pairs below): out of line for the
// lane-per-trial kernel - 32 trials share the instruction stream, and lanes that reach the same
// routine from different call sites RECONVERGE inside one out-of-line copy (3.9e8 -> 4.6e8
// events/s) - and inline for the warp-per-trial kernel, where one lane runs alone and a call is
// pure latency (2.8e8 -> 3.4e8).  Measured on B200, profiles/r01_harbor.md.
template <class State, bool IN_SHARED>
struct HarborSim {
    State *gst;                 // HBM-resident state (lane per trial)
    uint32_t warp_in_cta;       // which of the CTA's shared-memory states (warp per trial)

    // Every state access goes through here so that the compiler knows the address space: a
    // pointer kept in a struct member is generic, and generic loads / stores cost an address-range
    // check each (the on-chip kernel had 840 generic loads and 810 generic stores before this).
    __device__ __forceinline__ State &S() const
    {
        if (IN_SHARED) {
            return ((State *)harbor_smem)[warp_in_cta];
        }
        __builtin_assume(__isGlobal(gst));
        return *gst;
    }

    Sfc64 rng;
    const ZigHot *hot;
    double now, arr_mean, unload_small, sum_wait;
    uint64_t reactivated, through[2];
    SummaryAcc sys_time[2];

    __device__ uint32_t schedule(uint32_t act, uint32_t subj, double t)
    {
        const uint32_t key = S().fel.push(0u, t, 0, act, subj, (int32_t)SIG_SUCCESS);
        if (key == 0u) {
            S().status |= TRIAL_ERR_FEL_OVERFLOW;
        }
        return key;
    }

    __device__ __noinline__ void await_push_out(HarborProc &p, uint32_t type, uint32_t ref) { return await_push_impl(p, type, ref); }
    __device__ __forceinline__ void await_push(HarborProc &p, uint32_t type, uint32_t ref) { if (IN_SHARED && HARBOR_INLINE_ON_CHIP) return await_push_impl(p, type, ref); else return await_push_out(p, type, ref); }
    __device__ __forceinline__ void await_push_impl(HarborProc &p, uint32_t type, uint32_t ref)
    {
        if (p.n_awaits >= 2u) {
            S().status |= TRIAL_ERR_PROC_OVERFLOW;
            return;
        }
        p.await_type[1] = p.await_type[0];
        p.await_ref[1] = p.await_ref[0];
        p.await_type[0] = (uint8_t)type;
        p.await_ref[0] = ref;
        p.n_awaits++;
    }

    __device__ __noinline__ void await_remove_out(HarborProc &p, uint32_t type, bool any, uint32_t ref) { return await_remove_impl(p, type, any, ref); }
    __device__ __forceinline__ void await_remove(HarborProc &p, uint32_t type, bool any, uint32_t ref) { if (IN_SHARED && HARBOR_INLINE_ON_CHIP) return await_remove_impl(p, type, any, ref); else return await_remove_out(p, type, any, ref); }
    __device__ __forceinline__ void await_remove_impl(HarborProc &p, uint32_t type, bool any, uint32_t ref)
    {
        for (uint32_t k = 0u; k < p.n_awaits; k++) {
            if (p.await_type[k] == type && (any || p.await_ref[k] == ref)) {
                if (k == 0u) {
                    p.await_type[0] = p.await_type[1];
                    p.await_ref[0] = p.await_ref[1];
                }
                p.n_awaits--;
                return;
            }
        }
    }

    __device__ __noinline__ void hold_begin_out(uint32_t pid, double dur) { return hold_begin_impl(pid, dur); }
    __device__ __forceinline__ void hold_begin(uint32_t pid, double dur) { if (IN_SHARED && HARBOR_INLINE_ON_CHIP) return hold_begin_impl(pid, dur); else return hold_begin_out(pid, dur); }
    __device__ __forceinline__ void hold_begin_impl(uint32_t pid, double dur)        // src/cmb_process.c:262-273
    {
        HarborProc &p = S().proc[pid];
        p.hold_handle = schedule(ACT_WAKE_TIME, pid, now + dur);
        await_push(p, AWAIT_TIME, p.hold_handle);
    }

    template <class Guard>
    __device__ __noinline__ void wait_begin_out(Guard &g, uint32_t gid, uint32_t pid) { return wait_begin_impl(g, gid, pid); }
    template <class Guard>
    __device__ __forceinline__ void wait_begin(Guard &g, uint32_t gid, uint32_t pid) { if (IN_SHARED && HARBOR_INLINE_ON_CHIP) return wait_begin_impl(g, gid, pid); else return wait_begin_out(g, gid, pid); }
    template <class Guard>
    __device__ __forceinline__ void wait_begin_impl(Guard &g, uint32_t gid, uint32_t pid)    // src/cmb_resourceguard.c:125-152
    {
        HarborProc &p = S().proc[pid];
        p.guard_key = ++S().guard_seq;
        if (g.push(p.guard_key, now, 0, 0u, pid, 0) == 0u) {
            S().status |= TRIAL_ERR_GUARD_OVERFLOW;
        }
        await_push(p, AWAIT_RESOURCE, gid);
    }

    template <class Guard>
    __device__ void signal(Guard &g, bool demand_holds)                 // :202-226
    {
        if (g.count > 0u && demand_holds) {
            const uint32_t pid = g.slot[1].subj;
            g.pop();
            schedule(ACT_WAKE_RESOURCE, pid, now);
        }
    }

    __device__ void record(HarborPool &p)               // record_sample, src/cmb_resourcepool.c:239-247
    {
        time_weighted_sample(p.hist, (double)p.in_use, now);
    }

    // one pass of cmi_pool_acquire_inner's loop (no pre-emption): true when the claim is filled
    template <class Guard>
    __device__ __noinline__ bool pool_grab_out(HarborPool &p, Guard &g, uint8_t &rem, uint8_t &held) { return pool_grab_impl(p, g, rem, held); }
    template <class Guard>
    __device__ __forceinline__ bool pool_grab(HarborPool &p, Guard &g, uint8_t &rem, uint8_t &held) { if (IN_SHARED && HARBOR_INLINE_ON_CHIP) return pool_grab_impl(p, g, rem, held); else return pool_grab_out(p, g, rem, held); }
    template <class Guard>
    __device__ __forceinline__ bool pool_grab_impl(HarborPool &p, Guard &g, uint8_t &rem, uint8_t &held)
    {
        const uint32_t available = p.cap - p.in_use;
        if (available >= rem) {
            p.in_use += rem;
            record(p);
            held += rem;
            rem = 0u;
            signal(g, p.cap - p.in_use > 0u);
            return true;
        }
        if (available > 0u) {
            p.in_use += available;
            record(p);
            rem -= (uint8_t)available;
            held += (uint8_t)available;
        }
        return false;
    }

    template <class Guard>
    __device__ __noinline__ void pool_release_out(HarborPool &p, Guard &g, uint32_t amount, uint8_t &held) { return pool_release_impl(p, g, amount, held); }
    template <class Guard>
    __device__ __forceinline__ void pool_release(HarborPool &p, Guard &g, uint32_t amount, uint8_t &held) { if (IN_SHARED && HARBOR_INLINE_ON_CHIP) return pool_release_impl(p, g, amount, held); else return pool_release_out(p, g, amount, held); }
    template <class Guard>
    __device__ __forceinline__ void pool_release_impl(HarborPool &p, Guard &g, uint32_t amount, uint8_t &held)   // :561-605
    {
        held -= (uint8_t)amount;
        p.in_use -= amount;
        record(p);
        signal(g, p.cap - p.in_use > 0u);
    }

    __device__ bool can_dock(const HarborProc &s) const // is_ready_to_dock, test/test_condition.c:192-236
    {
        if (S().water_depth < (s.size == 0u ? 8.0 : 13.0)) return false;
        if (S().wind_magnitude > (s.size == 0u ? 10.0 : 12.0)) return false;
        if (S().tugs.cap - S().tugs.in_use < s.need) return false;
        return S().berth[s.size].cap - S().berth[s.size].in_use >= 1u;
    }

    // cmb_condition_signal, src/cmb_condition.c:120-167
    __device__ __noinline__ uint32_t harbormaster_signal_out() { return harbormaster_signal_impl(); }
    __device__ __forceinline__ uint32_t harbormaster_signal() { if (IN_SHARED && HARBOR_INLINE_ON_CHIP) return harbormaster_signal_impl(); else return harbormaster_signal_out(); }
    __device__ __forceinline__ uint32_t harbormaster_signal_impl()
    {
        auto &cv = S().harbormaster;
        uint32_t hit[State::GUARD];
        uint32_t cnt = 0u;
        for (uint32_t k = 1u; k <= cv.count; k++) {
            const uint32_t pid = cv.slot[k].subj;
            if (can_dock(S().proc[pid])) {
                hit[cnt++] = cv.slot[k].key;
                schedule(ACT_WAKE_CONDITION, pid, now);
            }
        }
        for (uint32_t k = 0u; k < cnt; k++) {
            cv.remove(hit[k]);
        }
        return cnt;
    }

    __device__ __noinline__ void davyjones_signal_out() { return davyjones_signal_impl(); }
    __device__ __forceinline__ void davyjones_signal() { if (IN_SHARED && HARBOR_INLINE_ON_CHIP) return davyjones_signal_impl(); else return davyjones_signal_out(); }
    __device__ __forceinline__ void davyjones_signal_impl()
    {
        auto &cv = S().davyjones;
        uint32_t hit[3];
        uint32_t cnt = 0u;
        for (uint32_t k = 1u; k <= cv.count; k++) {
            if (S().departed != HARBOR_NONE) {          // is_departed
                hit[cnt++] = cv.slot[k].key;
                schedule(ACT_WAKE_CONDITION, cv.slot[k].subj, now);
            }
        }
        for (uint32_t k = 0u; k < cnt; k++) {
            cv.remove(hit[k]);
        }
    }

    __device__ __noinline__ double pert_out(double min, double mode, double max) { return pert_impl(min, mode, max); }
    __device__ __forceinline__ double pert(double min, double mode, double max) { if (IN_SHARED && HARBOR_INLINE_ON_CHIP) return pert_impl(min, mode, max); else return pert_out(min, mode, max); }
    __device__ __forceinline__ double pert_impl(double min, double mode, double max)
    {
        return rnd_PERT_mod(rng, *hot, min, mode, max, 4.0);
    }

    __device__ void body(uint32_t pid);

    // cmb_process_stop, src/cmb_process.c:698-723
    __device__ __noinline__ void stop_out(uint32_t pid) { return stop_impl(pid); }
    __device__ __forceinline__ void stop(uint32_t pid) { if (IN_SHARED && HARBOR_INLINE_ON_CHIP) return stop_impl(pid); else return stop_out(pid); }
    __device__ __forceinline__ void stop_impl(uint32_t pid)
    {
        HarborProc &p = S().proc[pid];
        if (p.status != PROC_RUNNING) {
            return;
        }
        p.status = PROC_FINISHED;
        while (p.n_awaits > 0u) {                       // cmi_process_cancel_awaiteds, :581-620
            const uint32_t type = p.await_type[0];
            const uint32_t ref = p.await_ref[0];
            p.await_type[0] = p.await_type[1];
            p.await_ref[0] = p.await_ref[1];
            p.n_awaits--;
            if (type == AWAIT_TIME) {
                (void)S().fel.remove(ref);
            }
            // AWAIT_RESOURCE: looked up by process address, never found:
            // the guard entry stays behind and may swallow a later signal
        }
        uint32_t hit[8];
        uint32_t n = 0u;
        for (uint32_t k = 1u; k <= S().fel.count && n < 8u; k++) {     // cmb_event_pattern_cancel(ANY, p, ANY)
            if (S().fel.slot[k].subj == pid) {
                hit[n++] = S().fel.slot[k].key;
            }
        }
        for (uint32_t k = 0u; k < n; k++) {
            (void)S().fel.remove(hit[k]);
        }
        // cmi_process_drop_resources, :507-527: the tugs were listed last, so they go first
        if (p.held_tugs > 0u) {
            S().tugs.in_use -= p.held_tugs;
            p.held_tugs = 0u;
            signal(S().tug_guard, S().tugs.cap - S().tugs.in_use > 0u);
        }
        if (p.held_berth > 0u) {
            S().berth[p.size].in_use -= p.held_berth;   // berth guards never have waiters
            p.held_berth = 0u;
        }
    }
};

template <class State, bool IN_SHARED>
__device__ void HarborSim<State, IN_SHARED>::body(uint32_t pid)
{
    HarborProc &p = S().proc[pid];
    const uint32_t kind = pid < HARBOR_FIXED ? pid : HARBOR_FIXED;
    switch (kind * 10u + p.pc) {
    // ---- weather
    case 0:
        for (;;) {
            {
                const double gust = rnd_rayleigh(rng, *hot, 5.0);
                S().wind_magnitude = 0.5 * gust + 0.5 * S().wind_magnitude;
                const double d1 = pert(0.0, 225.0, 360.0);
                const double d2 = pert(0.0, 45.0, 360.0);
                S().wind_direction = 0.75 * d1 + 0.25 * d2;
                reactivated += harbormaster_signal();
            }
            hold_begin(pid, 1.0);
            p.pc = 1u;
            return;
    case 1:;
        }
    // ---- tide
    case 10:
        for (;;) {
            {
                const double half_month = 0.5 * 29.5 * 24.0;
                const double t = fmod(now, half_month);
                const double astro = 15.0 + 1.0 * harbor_sin(2.0 * M_PI * t / 12.4) + 0.5 * harbor_sin(2.0 * M_PI * t / 24.0)
                                   + 0.25 * harbor_sin(2.0 * M_PI * t / (0.5 * 29.5 * 24));
                const double surge = 0.5 * S().wind_magnitude
                                   - 0.5 * S().wind_magnitude * harbor_sin(S().wind_direction * M_PI / 180.0);
                S().water_depth = astro + surge;
                reactivated += harbormaster_signal();
            }
            hold_begin(pid, 1.0);
            p.pc = 1u;
            return;
    case 11:;
        }
    // ---- arrivals
    case 20:
        for (;;) {
            hold_begin(pid, gp_exponential(rng, *hot, arr_mean));
            p.pc = 1u;
            return;
    case 21:
            {
                uint32_t slot = HARBOR_FIXED;
                while (slot < State::PROCS && S().proc[slot].in_use) {
                    slot++;
                }
                const uint32_t id = ++S().next_id;
                const uint32_t size = rng.bernoulli(0.25);
                if (slot == State::PROCS) {
                    S().status |= TRIAL_ERR_PROC_OVERFLOW;      // the ship is lost: the trial is void from here on
                }
                else {
                    HarborProc &s = S().proc[slot];
                    s.pc = 0u;
                    s.status = PROC_CREATED;
                    s.n_awaits = 0u;
                    s.active = 0u;
                    s.in_use = 1u;
                    s.size = (uint8_t)size;
                    s.need = size == 0u ? 1u : 3u;
                    s.held_tugs = s.held_berth = s.rem = 0u;
                    s.id = id;
                    s.next_departed = HARBOR_NONE;
                    schedule(ACT_START, slot, now);
                }
            }
        }
    // ---- departures
    case 30:
        for (;;) {
            wait_begin(S().davyjones, 2u, pid);
            p.pc = 1u;
            return;
    case 31:
            await_remove(p, AWAIT_RESOURCE, false, 2u);
            {
                const uint32_t slot = S().departed;
                HarborProc &s = S().proc[slot];
                S().departed = s.next_departed;
                summary_add(sys_time[s.size], s.t_sys);
                sum_wait = sum_wait + s.t_sys;
                through[s.size] += 1u;
                s.in_use = 0u;                          // free(shp)
            }
        }
    // ---- dots
    case 40:
        for (;;) {
            hold_begin(pid, 24.0 * 7 * 52);
            p.pc = 1u;
            return;
    case 41:;
        }
    // ---- ships
    case 50:
        p.t_arr = now;
        p.active = 1u;
        if (++S().alive > S().most_alive) {
            S().most_alive = S().alive;
        }
        while (!can_dock(p)) {
            wait_begin(S().harbormaster, 0u, pid);
            p.pc = 1u;
            return;
    case 51:
            await_remove(p, AWAIT_RESOURCE, false, 0u);
        }
        p.rem = 1u;                                     // both are there: the predicate just said so
        (void)pool_grab(S().berth[p.size], S().berth_guard[p.size], p.rem, p.held_berth);
        p.rem = p.need;
        (void)pool_grab(S().tugs, S().tug_guard, p.rem, p.held_tugs);
        hold_begin(pid, pert(0.4, 0.5, 0.8));
        p.pc = 2u;
        return;
    case 52:
        pool_release(S().tugs, S().tug_guard, p.need, p.held_tugs);
        {
            const double tua = (p.size == 0u) ? unload_small : 1.5 * unload_small;
            hold_begin(pid, pert(0.75 * tua, tua, 2 * tua));
        }
        p.pc = 3u;
        return;
    case 53:
        p.rem = p.need;
        while (!pool_grab(S().tugs, S().tug_guard, p.rem, p.held_tugs)) {
            wait_begin(S().tug_guard, 1u, pid);
            p.pc = 4u;
            return;
    case 54:
            await_remove(p, AWAIT_RESOURCE, false, 1u);
        }
        hold_begin(pid, pert(0.4, 0.5, 0.8));
        p.pc = 5u;
        return;
    case 55:
        pool_release(S().berth[p.size], S().berth_guard[p.size], 1u, p.held_berth);
        pool_release(S().tugs, S().tug_guard, p.need, p.held_tugs);
        p.active = 0u;
        S().alive--;
        p.next_departed = S().departed;
        S().departed = (uint16_t)pid;
        davyjones_signal();
        p.t_sys = now - p.t_arr;
        p.status = PROC_FINISHED;                       // return -> cmb_process_exit: nothing held or awaited
        return;
    }
}

// one whole trial, start to finish, on the calling thread; `st` is in HBM (lane per trial) or in
// shared memory (warp per trial, run by the warp's first lane)
template <bool TRACE, class State, bool IN_SHARED>
__device__ void harbor_trial(const HarborArgs &a, uint64_t trial, State *global_state, uint32_t warp_in_cta,
                             const ZigHot *hot)
{
    __builtin_assume(__isShared(hot));
    HarborSim<State, IN_SHARED> s;
    s.gst = global_state;
    s.warp_in_cta = warp_in_cta;
    s.hot = hot;
    s.now = 0.0;
    s.sum_wait = 0.0;
    s.reactivated = 0u;
    s.through[0] = s.through[1] = 0u;
    s.sys_time[0] = summary_empty();
    s.sys_time[1] = summary_empty();
    s.arr_mean = a.arr_mean[trial];
    s.unload_small = a.unload_small[trial];
    s.rng.seed(fmix64(a.master_seed, a.first_trial + trial));

    s.S().fel.clear();
    s.S().harbormaster.clear();
    s.S().tug_guard.clear();
    s.S().davyjones.clear();
    s.S().berth_guard[0].clear();
    s.S().berth_guard[1].clear();
    s.S().wind_magnitude = s.S().wind_direction = s.S().water_depth = 0.0;
    s.S().guard_seq = 0u;
    s.S().status = TRIAL_OK;
    s.S().next_id = s.S().alive = s.S().most_alive = 0u;
    s.S().departed = HARBOR_NONE;
    s.S().tugs.cap = (uint32_t)a.tugs;
    s.S().berth[0].cap = 6u;
    s.S().berth[1].cap = 3u;
    s.S().tugs.in_use = s.S().berth[0].in_use = s.S().berth[1].in_use = 0u;
    for (uint32_t i = 0u; i < State::PROCS; i++) {
        HarborProc &p = s.S().proc[i];
        p.pc = 0u;
        p.status = PROC_CREATED;
        p.n_awaits = 0u;
        p.active = 0u;
        p.in_use = i < HARBOR_FIXED ? 1u : 0u;
        p.size = p.need = p.held_tugs = p.held_berth = p.rem = 0u;
        p.hold_handle = p.guard_key = p.id = 0u;
        p.next_departed = HARBOR_NONE;
    }

    // creation order of test/test_condition.c:523-586 fixes the event keys
    s.schedule(ACT_START, 0u, 0.0);
    s.schedule(ACT_START, 1u, 0.0);
    s.S().tugs.hist.start();
    s.S().berth[0].hist.start();
    s.S().berth[1].hist.start();
    s.record(s.S().tugs);                                 // cmb_resourcepool_start_recording
    s.record(s.S().berth[0]);
    s.record(s.S().berth[1]);
    s.schedule(ACT_START, 2u, 0.0);
    s.schedule(ACT_START, 3u, 0.0);
    s.schedule(ACT_USER, SUBJ_MODEL, (double)a.duration);
    s.schedule(ACT_START, 4u, 0.0);

    uint64_t pops = 0u;
    uint32_t deepest = 0u;
    for (;;) {
        deepest = max(deepest, s.S().fel.count);
        if (!s.S().fel.pop()) {
            break;
        }
        const HeapTag ev = s.S().fel.slot[0];
        s.now = ev.d;
        if (TRACE) {
            if (pops < a.trace_cap) {
                a.trace_key[trial * a.trace_cap + pops] = ev.key;
                a.trace_time[trial * a.trace_cap + pops] = s.now;
            }
        }
        pops++;
        const uint32_t pid = ev.subj;
        bool run = false;
        switch (ev.act) {
        case ACT_START:
            s.S().proc[pid].status = PROC_RUNNING;
            s.S().proc[pid].pc = 0u;
            run = true;
            break;
        case ACT_WAKE_TIME:
            s.await_remove(s.S().proc[pid], AWAIT_TIME, false, ev.key);
            run = true;
            break;
        case ACT_WAKE_RESOURCE:
            run = s.S().proc[pid].status == PROC_RUNNING;
            break;
        case ACT_WAKE_CONDITION:
            s.await_remove(s.S().proc[pid], AWAIT_RESOURCE, true, 0u);
            run = s.S().proc[pid].status == PROC_RUNNING;
            break;
        case ACT_USER:                                  // end_sim_evt, test/test_condition.c:462-484
            for (uint32_t i = 0u; i < HARBOR_FIXED; i++) {
                s.stop(i);
            }
            for (;;) {                                  // active ships in (arrival time, id) order = id order
                uint32_t first = State::PROCS;
                for (uint32_t i = HARBOR_FIXED; i < State::PROCS; i++) {
                    if (s.S().proc[i].in_use && s.S().proc[i].active &&
                        (first == State::PROCS || s.S().proc[i].id < s.S().proc[first].id)) {
                        first = i;
                    }
                }
                if (first == State::PROCS) {
                    break;
                }
                s.S().proc[first].active = 0u;
                s.stop(first);
            }
            break;
        }
        if (run) {
            s.body(pid);
        }
    }

    if (a.events)    a.events[trial] = pops;
    if (a.objects)   a.objects[trial] = s.through[0] + s.through[1];
    if (a.t_end)     a.t_end[trial] = s.now;
    if (a.sum_wait)  a.sum_wait[trial] = s.sum_wait;
    if (a.status)    a.status[trial] = s.S().status;
    if (a.max_queue) a.max_queue[trial] = s.S().most_alive;
    if (a.counters) {
        uint64_t *c = a.counters + trial * 8u;
        c[0] = s.through[0];
        c[1] = s.through[1];
        c[2] = (uint64_t)__double_as_longlong(s.sys_time[0].m1);
        c[3] = (uint64_t)__double_as_longlong(s.sys_time[1].m1);
        c[4] = s.S().tugs.hist.acc.count;
        c[5] = (uint64_t)__double_as_longlong(s.S().tugs.hist.acc.m1);
        c[6] = s.S().berth[0].hist.acc.count | (s.S().berth[1].hist.acc.count << 32);
        c[7] = s.reactivated;
    }
    (void)deepest;
}

// Lane per trial, state in HBM: 32 trials share a warp's instruction stream, each in its own
// process body - on this workload ~4 of 32 lanes execute any one instruction, and every state access
// is an L2 / HBM round trip (profiles/r01_harbor.md).
template <bool TRACE>
__global__ void __launch_bounds__(GUARDED_BLOCK)
harbor_kernel(const HarborArgs a)
{
    __shared__ ZigHot hot;
    stage_zig_hot(hot, true);
    __syncthreads();

    const uint64_t trial = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (trial >= a.num_trials) {
        return;
    }
    if (a.repair && a.status[trial] == TRIAL_OK) {
        return;                                         // the on-chip pass finished this one
    }
    harbor_trial<TRACE, HarborState, false>(a, trial, &((HarborState *)a.state)[trial], 0u, &hot);
}

// Warp per trial, state in SHARED memory (BASELINE.json north_star's mapping): for a model whose
// lanes would all be in different process bodies anyway, one trial per warp loses nothing to
// divergence, and its heaps, guards and process table answer in shared-memory latency instead of
// an L2 / HBM round trip.  The warp's first lane runs the trial (the scalar event loop); the
// CTA's warps are independent trials; persistent CTAs pull trials until none are left.
template <bool TRACE>
__global__ void __launch_bounds__(HARBOR_BLOCK_ON_CHIP)
harbor_on_chip_kernel(const HarborArgs a)
{
    __shared__ ZigHot hot;
    stage_zig_hot(hot, true);
    __syncthreads();

    if ((threadIdx.x & 31u) != 0u) {
        return;
    }
    constexpr uint32_t WARPS = HARBOR_BLOCK_ON_CHIP / 32;
    const uint32_t w = threadIdx.x >> 5;
    for (uint64_t trial = (uint64_t)blockIdx.x * WARPS + w; trial < a.num_trials; trial += (uint64_t)gridDim.x * WARPS) {
        harbor_trial<TRACE, HarborStateOnChip, true>(a, trial, nullptr, w, &hot);
    }
}

}  // namespace cimba_b200
