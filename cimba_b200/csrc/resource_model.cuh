// resource_model.cuh - test/test_resource.c as it stands (model 14): three "preemptable"
// processes with random priorities and one "preempter" (priority 0) on one cmb_resource whose
// usage history is recorded; an end event stops the four.
//
// cmb_resource_acquire / release / preempt with wakeup_event_preempt (src/cmb_resource.c:182-320),
// resource_drop_holder on stop (:45-57), record_sample (:107-118) folded into a time-weighted
// summary.  With duration 25 and seed 0x34f05c64d7ad598f the reference's own golden file
// test/reference/resource.txt says: history "N 30  Mean 0.9816", Target_3 pre-empted at t = 6.3280.
// guard[2] = the resource's guard, tool_holder = its holder (general.cuh).
#pragma once

#include "general.cuh"
#include "guarded_model.cuh"
#include "summary.cuh"

namespace cimba_b200 {

constexpr uint32_t RESOURCE_PROCS = 4u;            // 0..2 targets, 3 the preempter

__device__ __forceinline__ void resource_sample(GeneralSim &s, GuardedTally &t)
{
    time_weighted_sample(t.hist, s.st->tool_holder != NO_HOLDER ? 1.0 : 0.0, s.now);
}

__device__ void resource_body(GeneralSim &s, GuardedTally &t, uint32_t pid, int32_t sig)
{
    GeneralState *st = s.st;
    GenProc &p = st->proc[pid];
    switch (p.pc) {
    case 0:
        if (pid < 3u) {                                 // preemptable (test/test_resource.c:54-80)
            for (;;) {
                if (st->tool_holder == NO_HOLDER) {     // cmb_resource_acquire
                    st->tool_holder = pid;
                    p.holds_tool = 1u;
                    resource_sample(s, t);
                    sig = (int32_t)SIG_SUCCESS;
                }
                else {
                    s.wait_begin(2u, pid);
                    p.pc = 1u;
                    return;
    case 1:
                    sig = s.wait_end(2u, pid, sig);
                    if (sig == (int32_t)SIG_SUCCESS) {
                        st->tool_holder = pid;
                        p.holds_tool = 1u;
                        resource_sample(s, t);
                    }
                }
                if (sig == (int32_t)SIG_SUCCESS) {
                    t.c[0] += 1u;
                    p.stamp = s.now;
                    s.hold_begin(pid, gp_exponential(s.rng, *s.hot, 1.0));
                    p.pc = 2u;
                    return;
    case 2:
                    sig = s.hold_end(pid, sig);
                    if (sig == (int32_t)SIG_SUCCESS) {
                        p.holds_tool = 0u;              // cmb_resource_release
                        st->tool_holder = NO_HOLDER;
                        resource_sample(s, t);
                        s.signal(2u, true);
                        t.sum_wait = __dadd_rn(t.sum_wait, __dsub_rn(s.now, p.stamp));
                    }
                    else {                              // "someone stole it from me"
                        t.c[1] += 1u;
                        if (t.c[5] == 0u) {
                            t.c[4] = (uint64_t)__double_as_longlong(s.now);
                            t.c[5] = pid + 1u;
                        }
                    }
                }
                s.hold_begin(pid, gp_exponential(s.rng, *s.hot, 1.0));
                p.pc = 3u;
                return;
    case 3:
                (void)s.hold_end(pid, sig);
            }
        }
        for (;;) {                                      // preempter (:82-99)
            if (st->tool_holder == NO_HOLDER) {         // cmb_resource_preempt: free
                st->tool_holder = pid;
                p.holds_tool = 1u;
                resource_sample(s, t);
            }
            else if (p.prio >= st->proc[st->tool_holder].prio) {
                const uint32_t victim = st->tool_holder;        // kick it out: no sample, the resource stays occupied
                st->proc[victim].holds_tool = 0u;
                s.cancel_awaiteds(pid);                 // sic: the CALLER's awaiteds
                s.schedule(ACT_WAKE_PREEMPT, victim, (int32_t)SIG_PREEMPTED, s.now, st->proc[victim].prio);
                st->tool_holder = pid;
                p.holds_tool = 1u;
            }
            else {                                      // wait politely
                s.wait_begin(2u, pid);
                p.pc = 10u;
                return;
    case 10:
                sig = s.wait_end(2u, pid, sig);
                if (sig == (int32_t)SIG_SUCCESS) {
                    st->tool_holder = pid;
                    p.holds_tool = 1u;
                    resource_sample(s, t);
                }
            }
            t.c[2] += 1u;
            s.hold_begin(pid, gp_exponential(s.rng, *s.hot, 1.0));
            p.pc = 11u;
            return;
    case 11:
            (void)s.hold_end(pid, sig);
            p.holds_tool = 0u;
            st->tool_holder = NO_HOLDER;
            resource_sample(s, t);
            s.signal(2u, true);
            s.hold_begin(pid, gp_exponential(s.rng, *s.hot, 1.0));
            p.pc = 12u;
            return;
    case 12:
            (void)s.hold_end(pid, sig);
        }
    }
}

template <bool TRACE>
__global__ void __launch_bounds__(GUARDED_BLOCK)
resource_kernel(const GuardedArgs a)
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
    t.put_mean = t.get_mean = 1.0;
    t.record = 1u;
    t.use_pq = 0u;
    t.fillers = t.drainers = 0u;
    t.amount_max = 0;

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
    st->tool_observer = 0u;
    st->guard_seq = 0u;
    st->status = TRIAL_OK;
    st->ring_cap = 1u;
    st->ring_head = st->ring_len = 0u;
    st->n_ew = 0u;

    t.hist.start();
    t.hist.sample(0.0, 0.0);                            // cmb_resource_start_recording: idle at t = 0

    for (uint32_t i = 0u; i < RESOURCE_PROCS; i++) {
        GenProc &p = st->proc[i];
        p.pc = 0u;
        p.status = PROC_CREATED;
        p.kind = i;
        p.n_awaits = 0u;
        p.n_waiters = 0u;
        p.hold_handle = p.guard_key = 0u;
        p.stamp = 0.0;
        p.holds_pool = p.holds_tool = p.held = p.req = p.rem = p.initially_held = 0u;
        p.prio = (i < 3u) ? (int32_t)s.rng.dice(-5, 5) : 0;
        s.schedule(ACT_START, i, 0, s.now, p.prio);
    }
    s.schedule(ACT_USER, SUBJ_MODEL, 0, (double)a.duration, 0);

    uint64_t pops = 0u;
    for (;;) {
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
        case ACT_WAKE_PREEMPT:                          // src/cmb_resource.c:256-268
            run = st->proc[pid].status == PROC_RUNNING;
            break;
        case ACT_USER:                                  // end_sim_evt: stop the four; a holder drops the resource
            for (uint32_t i = 0u; i < RESOURCE_PROCS; i++) {
                const bool held = st->proc[i].status == PROC_RUNNING && st->proc[i].holds_tool != 0u;
                if (held) {
                    // resource_drop_holder samples between clearing the holder and signalling; the sample only
                    // reads the holder and the clock, so taking it first with the holder cleared is the same
                    const uint32_t keep = st->tool_holder;
                    st->tool_holder = NO_HOLDER;
                    resource_sample(s, t);
                    st->tool_holder = keep;
                }
                s.stop(i);
            }
            break;
        }
        if (run) {
            resource_body(s, t, pid, ev.arg);
        }
    }

    resource_sample(s, t);                              // cmb_resource_stop_recording
    t.c[3] = (uint64_t)__double_as_longlong(t.hist.acc.m1);
    if (a.events)    a.events[trial] = pops;
    if (a.objects)   a.objects[trial] = t.c[0] + t.c[2];
    if (a.t_end)     a.t_end[trial] = s.now;
    if (a.sum_wait)  a.sum_wait[trial] = t.sum_wait;
    if (a.status)    a.status[trial] = st->status;
    if (a.max_queue) a.max_queue[trial] = (uint32_t)t.hist.acc.count;
    if (a.counters) {
        for (int k = 0; k < 8; k++) {
            a.counters[trial * 8u + k] = t.c[k];
        }
    }
}

}  // namespace cimba_b200
