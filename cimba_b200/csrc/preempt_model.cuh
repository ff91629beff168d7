// preempt_model.cuh - resource pool with pre-emption (model 4).
//
// Workload: the reference's pool torture test, test/test_resourcepool.c:72-330, as
// restated with counters in oracle/ref_build/ref_driver.c (model 4): three mice that
// set their own priority and acquire politely, two rats that pre-empt, a cat that
// interrupts a random rodent, an end event that stops everyone.
//
// Parity vehicle for SURVEY.md section 8a rows a15 (cmi_pool_acquire_inner in full:
// greedy partial grabs, the holders heap, pre-emption of lower-priority holders,
// roll-back on interrupt, partial release, drop on stop), a6 (reprioritize) and a12
// (priority_set).  Built on general.cuh; each `case` is the continuation after a
// blocking call of the reference body.
#pragma once

#include "general.cuh"
#include "guarded_model.cuh"

namespace cimba_b200 {

constexpr uint32_t PREEMPT_MICE = 3u;
constexpr uint32_t PREEMPT_RODENTS = 5u;       // process PREEMPT_RODENTS is the cat

__device__ __forceinline__ void preempt_check(GeneralSim &s, GuardedTally &t, uint32_t pid)
{
    if (s.pool_held_by(pid) != s.st->proc[pid].held) {
        t.c[7] += 1u;
    }
}

__device__ __forceinline__ void preempt_take_signal(GeneralSim &s, GuardedTally &t, uint32_t pid, int32_t sig)
{
    if (sig == (int32_t)SIG_PREEMPTED) {
        t.c[2] += 1u;
        s.st->proc[pid].held = 0u;
    }
    else if (sig != (int32_t)SIG_SUCCESS) {
        t.c[3] += 1u;
    }
    t.c[4] += (uint64_t)(int64_t)sig;
}

__device__ void preempt_rodent(GeneralSim &s, GuardedTally &t, uint32_t pid, int32_t sig)
{
    GeneralState *st = s.st;
    GenProc &p = st->proc[pid];
    const bool rat = pid >= PREEMPT_MICE;
    switch (p.pc) {
    case 0:
        for (;;) {
            preempt_check(s, t, pid);
            p.req = (uint32_t)s.rng.dice(1, 5);
            if (!rat) {
                s.priority_set_self(pid, (int32_t)s.rng.dice(-5, 5));
            }
            // ---- cmi_pool_acquire_inner, src/cmb_resourcepool.c:362-533
            p.initially_held = s.pool_held_by(pid);
            p.rem = p.req;
            for (;;) {
                {
                    const uint32_t avail = st->pool_cap - st->pool_in_use;
                    if (avail >= p.rem) {
                        st->pool_in_use += p.rem;
                        s.pool_update_record(pid, p.rem);
                        s.pool_signal();
                        sig = (int32_t)SIG_SUCCESS;
                        goto acquired;
                    }
                    else if (avail > 0u) {
                        st->pool_in_use += avail;
                        p.rem -= avail;
                        s.pool_update_record(pid, avail);
                    }
                }
                if (rat) {
                    // mug holders of lower priority, lowest first (:430-478)
                    while (st->holders.count > 0u && st->holders.slot[1].prio < p.prio) {
                        st->holders.pop();
                        const uint32_t victim = st->holders.slot[0].subj;
                        const uint32_t loot = (uint32_t)st->holders.slot[0].arg;
                        st->proc[victim].holds_pool = 0u;               // cmi_process_remove_holdable
                        s.interrupt(victim, (int32_t)SIG_PREEMPTED, st->proc[victim].prio);
                        if (loot < p.rem) {
                            s.pool_update_record(pid, loot);
                            p.rem -= loot;
                        }
                        else {
                            s.pool_update_record(pid, p.rem);
                            st->pool_in_use -= loot - p.rem;
                            s.pool_signal();
                            sig = (int32_t)SIG_SUCCESS;
                            goto acquired;
                        }
                    }
                }
                s.wait_begin(0u, pid);
                p.pc = 1u;
                return;
    case 1:
                sig = s.wait_end(0u, pid, sig);
                if (sig == (int32_t)SIG_PREEMPTED) {
                    goto acquired;                      // thrown out instead: empty-handed (:483-488)
                }
                else if (sig != (int32_t)SIG_SUCCESS) {
                    if (p.initially_held > 0u) {        // roll back to the initial holding (:494-502)
                        const uint32_t k = st->holders.find(pid + 1u);
                        const uint32_t surplus = (uint32_t)st->holders.slot[k].arg - p.initially_held;
                        st->holders.slot[k].arg = (int32_t)p.initially_held;
                        st->pool_in_use -= surplus;
                        s.pool_signal();
                    }
                    else {                              // had nothing: put back all (:503-518)
                        st->pool_in_use -= s.pool_held_by(pid);
                        if (st->holders.remove(pid + 1u)) {
                            p.holds_pool = 0u;
                        }
                    }
                    goto acquired;
                }
            }
acquired:
            if (sig == (int32_t)SIG_SUCCESS) {
                p.held += p.req;
                t.c[rat ? 1 : 0] += 1u;
                preempt_check(s, t, pid);
                s.hold_begin(pid, gp_exponential(s.rng, *s.hot, 1.0));
                p.pc = 2u;
                return;
    case 2:
                sig = s.hold_end(pid, sig);
                if (sig == (int32_t)SIG_SUCCESS) {
                    uint32_t rel = (uint32_t)s.rng.dice(1, 5);
                    if (rel > p.held || s.rng.dice(0, 1) == 1) {
                        rel = p.held;
                    }
                    s.pool_release(pid, rel);
                    p.held -= rel;
                    t.c[5] += rel;
                    t.sum_wait = __dadd_rn(t.sum_wait, __dmul_rn(s.now, (double)rel));
                }
                else {
                    preempt_take_signal(s, t, pid, sig);
                }
            }
            else {
                preempt_take_signal(s, t, pid, sig);
            }
            preempt_check(s, t, pid);
            s.hold_begin(pid, gp_exponential(s.rng, *s.hot, 1.0));
            p.pc = 3u;
            return;
    case 3:
            sig = s.hold_end(pid, sig);
            if (sig != (int32_t)SIG_SUCCESS) {
                preempt_take_signal(s, t, pid, sig);
            }
        }
    }
}

__device__ void preempt_cat(GeneralSim &s, uint32_t pid, int32_t sig)
{
    GenProc &p = s.st->proc[pid];
    switch (p.pc) {
    case 0:
        for (;;) {
            s.hold_begin(pid, gp_exponential(s.rng, *s.hot, 1.0));
            p.pc = 1u;
            return;
    case 1:
            (void)s.hold_end(pid, sig);
            {
                const uint32_t victim = (uint32_t)s.rng.dice(0, PREEMPT_RODENTS - 1u);
                const int32_t loud = (int32_t)s.rng.dice(10, 100);
                const int32_t isig = (s.rng.dice(0, 1) == 1) ? (int32_t)SIG_INTERRUPTED : loud;
                s.interrupt(victim, isig, 0);
            }
        }
    }
}

template <bool TRACE>
__global__ void __launch_bounds__(GUARDED_BLOCK)
preempt_kernel(const GuardedArgs a)
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

    st->fel.clear();
    st->guard[0].clear();
    st->guard[1].clear();
    st->guard[2].clear();
    st->tool_holder = NO_HOLDER;
    st->buf_cap = st->buf_level = 0u;
    st->holders.clear();
    st->pool_cap = (uint32_t)a.capacity;
    st->pool_in_use = 0u;
    st->guard_seq = 0u;
    st->n_ew = 0u;
    st->tool_observer = 0u;
    st->status = TRIAL_OK;
    st->ring_cap = 1u;
    st->ring_head = st->ring_len = 0u;

    for (uint32_t i = 0u; i <= PREEMPT_RODENTS; i++) {
        GenProc &p = st->proc[i];
        p.pc = 0u;
        p.status = PROC_CREATED;
        p.kind = (i < PREEMPT_MICE) ? 0u : (i < PREEMPT_RODENTS) ? 1u : 2u;
        p.n_awaits = 0u;
        p.n_waiters = 0u;
        p.hold_handle = p.guard_key = 0u;
        p.stamp = 0.0;
        p.holds_pool = p.holds_tool = p.held = p.req = p.rem = p.initially_held = 0u;
        p.prio = (i < PREEMPT_RODENTS) ? (int32_t)s.rng.dice(-5, 5) : 0;
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
            run = st->proc[pid].status == PROC_RUNNING;
            break;
        case ACT_WAKE_INTERRUPT:
            s.cancel_awaiteds(pid);
            run = true;
            break;
        case ACT_USER:
            for (uint32_t i = 0u; i <= PREEMPT_RODENTS; i++) {
                s.stop(i);
            }
            break;
        }
        if (run) {
            if (pid == PREEMPT_RODENTS) {
                preempt_cat(s, pid, ev.arg);
            }
            else {
                preempt_rodent(s, t, pid, ev.arg);
            }
        }
    }

    t.c[6] = st->pool_in_use;
    if (a.events)    a.events[trial] = pops;
    if (a.objects)   a.objects[trial] = t.c[0] + t.c[1];
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
