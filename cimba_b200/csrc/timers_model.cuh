// timers_model.cuh - timers, process/event waits, resume/yield, exit + restart, event
// reschedule / reprioritize / cancel with waiter notification, guard observers (model 8).
//
// Workload: oracle/ref_build/ref_driver.c model 8, written against the reference API in
// the manner of tutorial/tut_3_1.c:370-430 (reneging: cmb_process_timer_add / timer_set /
// timers_clear around a blocking call, cmb_process_yield + cmb_process_resume),
// test/test_process.c:104-129 (cmb_process_wait_event, cmb_process_wait_process,
// cmb_process_exit) and test/test_event.c:175-181 (cmb_event_reschedule / reprioritize).
// It closes SURVEY.md section 8a rows a9 (reschedule, reprioritize, event waiters), a11
// (timers as separate calls), a12 (wait_process, wait_event, resume, exit) and a13
// (observers: src/cmb_resourceguard.c:227-239 forwards every signal of the desk's guard
// to the condition's guard).
//
// Processes (indices fixed by creation order, as in the reference driver):
//   0,1 patients  2 clerk  3 supervisor  4 ringer  5 listener  6 watcher  7 nuisance
// guard[2] = the desk's guard (cmb_resource), guard[0] = the condition's guard.
#pragma once

#include "general.cuh"
#include "guarded_model.cuh"

namespace cimba_b200 {

constexpr uint32_t TIMERS_PROCS = 8u;
constexpr uint32_t TIMERS_CLERK = 2u;
constexpr int32_t  TIMERS_SIG_ALARM = 77;
constexpr int32_t  TIMERS_SIG_DOZE = 55;
constexpr int32_t  TIMERS_SIG_NUDGE = 9;

__device__ __forceinline__ void timers_note(GuardedTally &t, int32_t sig)
{
    if (sig != (int32_t)SIG_SUCCESS) {
        t.c[7] += (uint64_t)(int64_t)sig;
    }
}

// the eight process bodies as resume points; kind = process role, pc = continuation
__device__ void timers_body(GeneralSim &s, GuardedTally &t, uint32_t pid, int32_t sig)
{
    GeneralState *st = s.st;
    GenProc &p = st->proc[pid];
    switch (p.kind * 10u + p.pc) {
    // ---- patients (t_patient_body)
    case 0:
        for (;;) {
            s.hold_begin(pid, gp_exponential(s.rng, *s.hot, t.put_mean));
            p.pc = 1u;
            return;
    case 1:
            timers_note(t, s.hold_end(pid, sig));
            p.timer = s.timer_add(pid, gp_exponential(s.rng, *s.hot, __dmul_rn(2.0, t.get_mean)), (int32_t)SIG_TIMEOUT);
            if (st->tool_holder == NO_HOLDER) {         // cmb_resource_acquire, src/cmb_resource.c:191-229
                st->tool_holder = pid;
                p.holds_tool = 1u;
                sig = (int32_t)SIG_SUCCESS;
            }
            else {
                s.wait_begin(2u, pid);
                p.pc = 2u;
                return;
    case 2:
                sig = s.wait_end(2u, pid, sig);
                if (sig == (int32_t)SIG_SUCCESS) {
                    st->tool_holder = pid;
                    p.holds_tool = 1u;
                }
            }
            if (sig == (int32_t)SIG_SUCCESS) {
                (void)s.timer_cancel(pid, p.timer);
                t.c[0] += 1u;
                p.stamp = s.now;
                (void)s.timer_add(pid, gp_exponential(s.rng, *s.hot, 3.0), TIMERS_SIG_ALARM);
                s.hold_begin(pid, gp_exponential(s.rng, *s.hot, t.get_mean));
                p.pc = 3u;
                return;
    case 3:
                timers_note(t, s.hold_end(pid, sig));
                s.timers_clear(pid);
                p.holds_tool = 0u;                      // cmb_resource_release, :234-250
                st->tool_holder = NO_HOLDER;
                s.tool_signal();
                t.sum_wait = __dadd_rn(t.sum_wait, __dsub_rn(s.now, p.stamp));
                s.timers_clear(pid);                    // cmb_process_timer_set = clear + add
                (void)s.timer_add(pid, gp_exponential(s.rng, *s.hot, 0.3), TIMERS_SIG_DOZE);
                p.pc = 4u;                              // cmb_process_yield
                return;
    case 4:
                timers_note(t, sig);
                if (sig != TIMERS_SIG_DOZE) {
                    s.timers_clear(pid);
                }
            }
            else if (sig == (int32_t)SIG_TIMEOUT) {
                t.c[1] += 1u;
            }
            else {
                timers_note(t, sig);
                s.timers_clear(pid);
            }
        }
    // ---- clerk (t_clerk_body)
    case 10:
        st->clerk_start_pending = 0u;
        p.jobs = (int32_t)s.rng.dice(2, 5);
        for (p.job = 0; p.job < p.jobs; p.job++) {
            s.hold_begin(pid, gp_exponential(s.rng, *s.hot, 1.0));
            p.pc = 1u;
            return;
    case 11:
            timers_note(t, s.hold_end(pid, sig));
            if (s.rng.dice(0, 2) == 0) {
                s.resume((uint32_t)s.rng.dice(0, 1), TIMERS_SIG_NUDGE);
            }
            t.c[3] += 1u;
        }
        s.exit(pid);
        return;
    // ---- supervisor (t_supervisor_body)
    case 20:
        for (;;) {
            if (st->proc[TIMERS_CLERK].status == PROC_FINISHED) {       // src/cmb_process.c:428-452
                sig = (int32_t)SIG_SUCCESS;
            }
            else {
                s.wait_process_begin(pid, TIMERS_CLERK);
                p.pc = 1u;
                return;
            }
            // fall through
    case 21:
            if (sig == (int32_t)SIG_SUCCESS) {
                t.c[2] += 1u;
                s.hold_begin(pid, gp_exponential(s.rng, *s.hot, 0.5));
                p.pc = 2u;
                return;
    case 22:
                timers_note(t, s.hold_end(pid, sig));
                if (st->proc[TIMERS_CLERK].status == PROC_FINISHED && st->clerk_start_pending == 0u) {
                    st->clerk_start_pending = 1u;
                    s.schedule(ACT_START, TIMERS_CLERK, 0, s.now, st->proc[TIMERS_CLERK].prio);
                }
            }
            else {
                timers_note(t, sig);
            }
        }
    // ---- ringer (t_ringer_body)
    case 30:
        for (;;) {
            {
                const double when = __dadd_rn(s.now, gp_exponential(s.rng, *s.hot, 2.0));
                const int32_t pri = (int32_t)s.rng.dice(-2, 2);
                p.bell = s.schedule(ACT_BELL, SUBJ_MODEL, 0, when, pri);
                st->bell = p.bell;
            }
            s.hold_begin(pid, gp_exponential(s.rng, *s.hot, 0.7));
            p.pc = 1u;
            return;
    case 31:
            timers_note(t, s.hold_end(pid, sig));
            if (s.event_is_scheduled(p.bell)) {
                const long long op = s.rng.dice(0, 3);
                if (op == 0) {
                    (void)s.event_reschedule(p.bell, __dadd_rn(s.now, gp_exponential(s.rng, *s.hot, 1.0)));
                    t.c[5] += 1u;
                }
                else if (op == 1) {
                    (void)s.event_reprioritize(p.bell, (int32_t)s.rng.dice(-5, 5));
                    t.c[5] += 100u;
                }
                else if (op == 2) {
                    (void)s.event_cancel(p.bell);
                    t.c[5] += 10000u;
                }
            }
            if (s.event_is_scheduled(p.bell)) {
                s.wait_event_begin(pid, p.bell);
                p.pc = 2u;
                return;
    case 32:
                timers_note(t, sig);
            }
        }
    // ---- listener (t_listener_body)
    case 40:
        for (;;) {
            p.bell = st->bell;
            if (p.bell != 0u && s.event_is_scheduled(p.bell)) {
                s.wait_event_begin(pid, p.bell);
                p.pc = 1u;
                return;
    case 41:
                if (sig == (int32_t)SIG_SUCCESS) {
                    t.c[6] += 1000u;
                }
                else {
                    timers_note(t, sig);
                }
            }
            else {
                s.hold_begin(pid, gp_exponential(s.rng, *s.hot, 0.5));
                p.pc = 2u;
                return;
    case 42:
                timers_note(t, s.hold_end(pid, sig));
            }
        }
    // ---- watcher (t_watcher_body): cmb_condition_wait = cmb_resourceguard_wait on the condition's guard
    case 50:
        for (;;) {
            s.wait_begin(0u, pid);
            p.pc = 1u;
            return;
    case 51:
            sig = s.wait_end(0u, pid, sig);
            if (sig == (int32_t)SIG_SUCCESS) {
                t.c[6] += 1u;
            }
            else {
                timers_note(t, sig);
            }
            s.hold_begin(pid, gp_exponential(s.rng, *s.hot, 0.8));
            p.pc = 2u;
            return;
    case 52:
            timers_note(t, s.hold_end(pid, sig));
        }
    // ---- nuisance (t_nuisance_body)
    case 60:
        for (;;) {
            s.hold_begin(pid, gp_exponential(s.rng, *s.hot, 1.0));
            p.pc = 1u;
            return;
    case 61:
            (void)s.hold_end(pid, sig);
            {
                const uint32_t victim = (uint32_t)s.rng.dice(0, TIMERS_PROCS - 2u);
                const int32_t isig = (int32_t)s.rng.dice(1, 10);
                const int32_t ipri = (int32_t)s.rng.dice(-5, 5);
                if (st->proc[victim].status == PROC_RUNNING) {
                    s.interrupt(victim, isig, ipri);
                }
            }
        }
    }
}

template <bool TRACE>
__global__ void __launch_bounds__(GUARDED_BLOCK)
timers_kernel(const GuardedArgs a)
{
    __shared__ ZigHot hot;
    stage_zig_hot(hot, false);
    __syncthreads();

    const uint64_t trial = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (trial >= a.num_trials) {
        return;
    }
    GeneralState *st = &a.state[trial];
    GeneralSim s;
    s.st = st;
    s.hot = &hot;
    s.now = 0.0;
    s.rng.seed(fmix64(a.master_seed, a.first_trial + trial));

    GuardedTally t;
    for (int k = 0; k < 8; k++) {
        t.c[k] = 0u;
    }
    t.sum_wait = 0.0;
    t.put_mean = a.put_mean[trial];                     // arr_mean
    t.get_mean = a.get_mean[trial];                     // srv_mean

    st->fel.clear();
    st->guard[0].clear();
    st->guard[1].clear();
    st->guard[2].clear();
    st->holders.clear();
    st->pq.clear();
    st->pq_cap = 0u;
    st->pool_cap = st->pool_in_use = 0u;
    st->buf_cap = st->buf_level = 0u;
    st->tool_holder = NO_HOLDER;
    st->tool_observer = 1u;                             // cmb_resourceguard_register(desk guard, condition guard)
    st->guard_seq = 0u;
    st->status = TRIAL_OK;
    st->ring_cap = 1u;
    st->ring_head = st->ring_len = 0u;
    st->n_ew = 0u;
    st->bell = 0u;
    st->clerk_start_pending = 0u;

    const uint32_t kind_of[TIMERS_PROCS] = { 0u, 0u, 1u, 2u, 3u, 4u, 5u, 6u };
    for (uint32_t i = 0u; i < TIMERS_PROCS; i++) {
        GenProc &p = st->proc[i];
        p.pc = 0u;
        p.status = PROC_CREATED;
        p.kind = kind_of[i];
        p.n_awaits = 0u;
        p.n_waiters = 0u;
        p.hold_handle = p.guard_key = 0u;
        p.stamp = 0.0;
        p.holds_pool = p.holds_tool = p.held = p.req = p.rem = p.initially_held = 0u;
        p.timer = p.bell = 0u;
        p.jobs = p.job = 0;
        p.prio = (i + 1u < TIMERS_PROCS) ? (int32_t)s.rng.dice(-5, 5) : 0;
        s.schedule(ACT_START, i, 0, s.now, p.prio);
    }
    s.schedule(ACT_USER, SUBJ_MODEL, 0, (double)a.duration, 0);

    uint64_t pops = 0u;
    uint32_t deepest = 0u;
    for (;;) {
        deepest = max(deepest, st->fel.count);
        if (!st->fel.pop()) {
            break;
        }
        const HeapTag ev = st->fel.slot[0];
        s.now = ev.d;
        if (TRACE) {
            if (pops < a.trace_cap) {
                a.trace_key[trial * a.trace_cap + pops] = ev.key;
                a.trace_time[trial * a.trace_cap + pops] = s.now;
            }
        }
        pops++;
        if (st->n_ew != 0u) {                           // src/cmb_event.c:243-246
            s.wake_event_waiters(ev.key, (int32_t)SIG_SUCCESS);
        }
        const uint32_t pid = ev.subj;
        bool run = false;
        switch (ev.act) {
        case ACT_START:
            st->proc[pid].status = PROC_RUNNING;
            st->proc[pid].pc = 0u;
            run = true;
            break;
        case ACT_WAKE_TIME:
            (void)s.await_remove(st->proc[pid], AWAIT_TIME, ev.key);
            run = true;
            break;
        case ACT_WAKE_RESOURCE:
            run = st->proc[pid].status == PROC_RUNNING;
            break;
        case ACT_WAKE_PROCESS:
            (void)s.await_remove_any(st->proc[pid], AWAIT_PROCESS);
            run = st->proc[pid].status == PROC_RUNNING;
            break;
        case ACT_WAKE_EVENT:
            (void)s.await_remove_any(st->proc[pid], AWAIT_EVENT);
            run = st->proc[pid].status == PROC_RUNNING;
            break;
        case ACT_RESUME:
            run = true;
            break;
        case ACT_WAKE_INTERRUPT:
            s.cancel_awaiteds(pid);
            run = true;
            break;
        case ACT_BELL:
            t.c[4] += 1u;
            break;
        case ACT_USER:
            for (uint32_t i = 0u; i < TIMERS_PROCS; i++) {
                s.stop(i);
            }
            break;
        }
        if (run) {
            timers_body(s, t, pid, ev.arg);
        }
    }

    if (a.events)    a.events[trial] = pops;
    if (a.objects)   a.objects[trial] = t.c[0];
    if (a.t_end)     a.t_end[trial] = s.now;
    if (a.sum_wait)  a.sum_wait[trial] = t.sum_wait;
    if (a.status)    a.status[trial] = st->status;
    if (a.max_queue) a.max_queue[trial] = deepest;
    if (a.counters) {
        for (int k = 0; k < 8; k++) {
            a.counters[trial * 8u + k] = t.c[k];
        }
    }
}

}  // namespace cimba_b200
