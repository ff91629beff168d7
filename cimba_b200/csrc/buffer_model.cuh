// buffer_model.cuh - cmb_buffer with partial fulfilment + binary cmb_resource with
// pre-emption (model 5).
//
// Workload: oracle/ref_build/ref_driver.c model 5, in the manner of the reference's
// test/test_buffer.c and test/test_resource.c: two fillers and two drainers moving
// random amounts through a buffer of finite capacity, a polite and a pre-empting
// worker sharing one tool, a nuisance interrupting all six, an end event.
//
// Parity vehicle for SURVEY.md section 8a row a16:
//   cmb_buffer_get / cmb_buffer_put   src/cmb_buffer.c:194-264, 279-346 (take/put what
//       fits, signal the opposite guard after every level change, re-signal the same
//       side while leftovers remain, signal once more before waiting, return the
//       partial amount on interrupt)
//   cmb_resource_acquire/release/preempt   src/cmb_resource.c:182-320 (kick the holder
//       out if my priority >= its priority; the kicked holder is resumed by
//       wakeup_event_preempt WITHOUT its awaiteds being cancelled; the CALLER's are)
#pragma once

#include "general.cuh"
#include "guarded_model.cuh"

namespace cimba_b200 {

constexpr uint32_t BUFFER_PROCS = 6u;          // 0,1 fillers; 2,3 drainers; 4 polite worker; 5 pushy worker; 6 nuisance

__device__ __forceinline__ void buffer_note(GuardedTally &t, int32_t sig)
{
    if (sig != (int32_t)SIG_SUCCESS) {
        t.c[6] += (uint64_t)(int64_t)sig;
    }
}

__device__ void buffer_body(GeneralSim &s, GuardedTally &t, uint32_t pid, int32_t sig)
{
    GeneralState *st = s.st;
    GenProc &p = st->proc[pid];
    switch (p.pc) {
    case 0:
        if (pid == BUFFER_PROCS) {                      // nuisance
            for (;;) {
                s.hold_begin(pid, gp_exponential(s.rng, *s.hot, 1.0));
                p.pc = 10u;
                return;
    case 10:
                (void)s.hold_end(pid, sig);
                {
                    const uint32_t victim = (uint32_t)s.rng.dice(0, BUFFER_PROCS - 1u);
                    const int32_t isig = (int32_t)s.rng.dice(1, 10);
                    const int32_t ipri = (int32_t)s.rng.dice(-5, 5);
                    s.interrupt(victim, isig, ipri);
                }
            }
        }
        if (pid < t.fillers) {                          // filler: cmb_buffer_put
            for (;;) {
                s.hold_begin(pid, gp_exponential(s.rng, *s.hot, t.put_mean));
                p.pc = 20u;
                return;
    case 20:
                buffer_note(t, s.hold_end(pid, sig));
                p.req = (uint32_t)s.rng.dice(1, t.amount_max);
                p.rem = p.req;
                for (;;) {
                    if (st->buf_cap - st->buf_level >= p.rem) {
                        st->buf_level += p.rem;
                        if (t.record) time_weighted_sample(t.hist, (double)st->buf_level, s.now);   // record_sample, src/cmb_buffer.c:129-136
                        p.rem = 0u;
                        s.signal(0u, st->buf_level > 0u);
                        if (st->buf_level < st->buf_cap) {
                            s.signal(1u, st->buf_level < st->buf_cap);
                        }
                        sig = (int32_t)SIG_SUCCESS;
                        break;
                    }
                    else if (st->buf_level < st->buf_cap) {
                        const uint32_t grab = st->buf_cap - st->buf_level;
                        st->buf_level = st->buf_cap;
                        if (t.record) time_weighted_sample(t.hist, (double)st->buf_level, s.now);   // record_sample, src/cmb_buffer.c:129-136
                        p.rem -= grab;
                        s.signal(0u, st->buf_level > 0u);
                    }
                    s.signal(0u, st->buf_level > 0u);
                    s.wait_begin(1u, pid);
                    p.pc = 21u;
                    return;
    case 21:
                    sig = s.wait_end(1u, pid, sig);
                    if (sig != (int32_t)SIG_SUCCESS) {
                        break;
                    }
                }
                t.c[0] += p.req - p.rem;
                if (sig != (int32_t)SIG_SUCCESS) {
                    t.c[2] += 1u;
                    buffer_note(t, sig);
                }
            }
        }
        if (pid < t.fillers + t.drainers) {             // drainer: cmb_buffer_get
            for (;;) {
                s.hold_begin(pid, gp_exponential(s.rng, *s.hot, t.get_mean));
                p.pc = 30u;
                return;
    case 30:
                buffer_note(t, s.hold_end(pid, sig));
                p.rem = (uint32_t)s.rng.dice(1, t.amount_max);
                p.held = 0u;                            // amount obtained so far
                for (;;) {
                    if (st->buf_level >= p.rem) {
                        st->buf_level -= p.rem;
                        if (t.record) time_weighted_sample(t.hist, (double)st->buf_level, s.now);   // record_sample, src/cmb_buffer.c:129-136
                        p.held += p.rem;
                        s.signal(1u, st->buf_level < st->buf_cap);
                        if (st->buf_level > 0u) {
                            s.signal(0u, st->buf_level > 0u);
                        }
                        sig = (int32_t)SIG_SUCCESS;
                        break;
                    }
                    else if (st->buf_level > 0u) {
                        const uint32_t grab = st->buf_level;
                        st->buf_level = 0u;
                        if (t.record) time_weighted_sample(t.hist, (double)st->buf_level, s.now);   // record_sample, src/cmb_buffer.c:129-136
                        p.held += grab;
                        p.rem -= grab;
                        s.signal(1u, st->buf_level < st->buf_cap);
                    }
                    s.signal(1u, st->buf_level < st->buf_cap);
                    s.wait_begin(0u, pid);
                    p.pc = 31u;
                    return;
    case 31:
                    sig = s.wait_end(0u, pid, sig);
                    if (sig != (int32_t)SIG_SUCCESS) {
                        break;
                    }
                }
                t.c[1] += p.held;
                if (sig != (int32_t)SIG_SUCCESS) {
                    t.c[3] += 1u;
                    buffer_note(t, sig);
                }
            }
        }
        for (;;) {                                      // workers: cmb_resource
            if (pid == 5u && st->tool_holder != NO_HOLDER && p.prio >= st->proc[st->tool_holder].prio) {
                // cmb_resource_preempt, kick-out branch (src/cmb_resource.c:282-299)
                const uint32_t victim = st->tool_holder;
                st->proc[victim].holds_tool = 0u;
                s.cancel_awaiteds(pid);                 // sic: the CALLER's awaiteds
                st->tool_holder = NO_HOLDER;
                s.schedule(ACT_WAKE_PREEMPT, victim, (int32_t)SIG_PREEMPTED, s.now, st->proc[victim].prio);
                st->tool_holder = pid;
                p.holds_tool = 1u;
                sig = (int32_t)SIG_SUCCESS;
            }
            else if (st->tool_holder == NO_HOLDER) {
                st->tool_holder = pid;
                p.holds_tool = 1u;
                sig = (int32_t)SIG_SUCCESS;
            }
            else {
                s.wait_begin(2u, pid);
                p.pc = 40u;
                return;
    case 40:
                sig = s.wait_end(2u, pid, sig);
                if (sig == (int32_t)SIG_SUCCESS) {
                    st->tool_holder = pid;
                    p.holds_tool = 1u;
                }
            }
            if (sig == (int32_t)SIG_SUCCESS) {
                t.c[4] += 1u;
                p.stamp = s.now;
                s.hold_begin(pid, gp_exponential(s.rng, *s.hot, 1.0));
                p.pc = 41u;
                return;
    case 41:
                sig = s.hold_end(pid, sig);
                if (sig == (int32_t)SIG_PREEMPTED) {
                    t.c[5] += 1u;
                    buffer_note(t, sig);
                }
                else {
                    buffer_note(t, sig);
                    p.holds_tool = 0u;                  // cmb_resource_release, :234-250
                    st->tool_holder = NO_HOLDER;
                    s.signal(2u, true);
                    t.sum_wait = __dadd_rn(t.sum_wait, __dsub_rn(s.now, p.stamp));
                }
            }
            else {
                buffer_note(t, sig);
            }
            s.hold_begin(pid, gp_exponential(s.rng, *s.hot, 1.0));
            p.pc = 42u;
            return;
    case 42:
            buffer_note(t, s.hold_end(pid, sig));
        }
    }
}

template <bool TRACE>
__global__ void __launch_bounds__(GUARDED_BLOCK)
buffer_kernel(const GuardedArgs a)
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
    t.put_mean = a.put_mean[trial];
    t.get_mean = a.get_mean[trial];
    t.record = a.record;                                // model 12 = test/test_buffer.c as it stands
    t.fillers = a.record ? 3u : 2u;
    t.drainers = a.record ? 3u : 2u;
    t.amount_max = a.record ? 15 : 8;
    t.hist.start();
    if (t.record) {
        t.hist.sample(0.0, 0.0);                        // cmb_buffer_recording_start: level 0 at t = 0
    }

    st->fel.clear();
    st->guard[0].clear();
    st->guard[1].clear();
    st->guard[2].clear();
    st->holders.clear();
    st->pool_cap = st->pool_in_use = 0u;
    st->buf_cap = (uint32_t)a.capacity;
    st->buf_level = 0u;
    st->tool_holder = NO_HOLDER;
    st->guard_seq = 0u;
    st->n_ew = 0u;
    st->tool_observer = 0u;
    st->status = TRIAL_OK;
    st->ring_cap = 1u;
    st->ring_head = st->ring_len = 0u;

    for (uint32_t i = 0u; i <= BUFFER_PROCS; i++) {
        GenProc &p = st->proc[i];
        p.pc = 0u;
        p.status = PROC_CREATED;
        p.kind = i;
        p.n_awaits = 0u;
        p.n_waiters = 0u;
        p.hold_handle = p.guard_key = 0u;
        p.stamp = 0.0;
        p.holds_pool = p.holds_tool = p.held = p.req = p.rem = p.initially_held = 0u;
        p.prio = (i < BUFFER_PROCS) ? (int32_t)s.rng.dice(-5, 5) : 0;
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
        case ACT_WAKE_PREEMPT:
            run = st->proc[pid].status == PROC_RUNNING;
            break;
        case ACT_WAKE_INTERRUPT:
            s.cancel_awaiteds(pid);
            run = true;
            break;
        case ACT_USER:
            for (uint32_t i = 0u; i <= BUFFER_PROCS; i++) {
                s.stop(i);
            }
            break;
        }
        if (run) {
            buffer_body(s, t, pid, ev.arg);
        }
    }

    t.c[7] = st->buf_level;
    if (t.record) {                                     // recording_stop + cmb_timeseries_summarize
        t.hist.sample((double)st->buf_level, s.now);
        t.c[4] = (uint64_t)__double_as_longlong(t.hist.acc.m1);
        deepest = (uint32_t)t.hist.acc.count;
    }
    if (a.events)    a.events[trial] = pops;
    if (a.objects)   a.objects[trial] = t.c[1];
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
