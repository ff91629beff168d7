// guarded_model.cuh - bounded object queue under interrupts (model 3).
//
// Workload: the reference's object-queue torture test, test/test_objectqueue.c:40-170,
// as restated with counters in oracle/ref_build/ref_driver.c (model 3): three
// putters and three getters with random priorities on a cmb_objectqueue of finite
// capacity (front AND rear guard in play), a nuisance process that interrupts a
// random worker with a random signal at a random event priority, and an end event
// at t = duration that stops every process.
//
// It is the parity vehicle for the cancel / interrupt / stop path of SURVEY.md
// section 8a rows a6, a9, a12 and the interrupted branches of a13/a14: event cancel
// by handle, pattern cancel by subject, guard self-removal by key, awaitable
// lists, unequal process and event priorities in both orders.
//
// Process bodies are resume-point machines over the GeneralSim primitives; each
// `case` is the continuation after a blocking call of the reference body.
#pragma once

#include "general.cuh"
#include "summary.cuh"

namespace cimba_b200 {

constexpr int GUARDED_BLOCK = 64;
constexpr uint32_t GUARDED_WORKERS = 6u;
constexpr uint32_t GUARDED_PUTTERS = 3u;
constexpr uint32_t GUARDED_NUISANCE = 6u;      // process index of the nuisance
constexpr uint32_t SUBJ_MODEL = 0xffffu;

struct GuardedArgs {
    int32_t  capacity;
    uint64_t master_seed, first_trial, num_trials, duration;
    const double *put_mean, *get_mean;
    uint64_t *events, *objects;
    double   *t_end, *sum_wait;
    uint32_t *status, *max_queue;
    uint64_t *counters;                // [num_trials][8]
    GeneralState *state;               // [num_trials]
    uint32_t  record;                  // models 11-13: fold the history into a time-weighted summary
    uint32_t  use_pq;                  // model 13: the objects sit in a cmb_priorityqueue
    uint64_t  trace_cap;
    uint64_t *trace_key;
    double   *trace_time;
};

struct GuardedTally {
    uint64_t c[8];
    double   sum_wait;
    double   put_mean, get_mean;
    uint32_t record;
    TimeWeighted hist;                 // cmb_objectqueue_recording_start (test/test_objectqueue.c:191)
    uint32_t use_pq;                   // model 13 = test/test_priorityqueue.c: priority desc, then FIFO
    uint32_t fillers, drainers;        // buffer models: 2 + 2 (model 5) or 3 + 3 (model 12 = test/test_buffer.c)
    int32_t  amount_max;               // ... moving 1..8 or 1..15 units
};

__device__ __forceinline__ void guarded_note(GuardedTally &t, int32_t sig, int which)
{
    if (sig != (int32_t)SIG_SUCCESS) {
        t.c[which] += 1u;
        t.c[5] += (uint64_t)(int64_t)sig;
    }
}

// putter / getter / nuisance bodies (ref_driver.c g_putter_body, g_getter_body,
// g_nuisance_body).  `sig` is the signal the process is resumed with.
__device__ void guarded_body(GeneralSim &s, GuardedTally &t, uint32_t pid, int32_t sig)
{
    GeneralState *st = s.st;
    GenProc &p = st->proc[pid];
    switch (p.pc) {
    case 0:
        for (;;) {
            {
                const double mean = (p.kind == 0u) ? t.put_mean : (p.kind == 1u) ? t.get_mean : 1.0;
                s.hold_begin(pid, gp_exponential(s.rng, *s.hot, mean));
            }
            p.pc = 1u;
            return;
    case 1:
            sig = s.hold_end(pid, sig);
            if (p.kind == 2u) {
                const uint32_t victim = (uint32_t)s.rng.dice(0, GUARDED_WORKERS - 1u);
                const int32_t isig = (int32_t)s.rng.dice(1, 10);
                const int32_t ipri = (int32_t)s.rng.dice(-5, 5);
                t.c[7] += 1u;
                s.interrupt(victim, isig, ipri);
                continue;
            }
            guarded_note(t, sig, 2);
            if (p.kind == 0u) {
                p.stamp = s.now;
                for (;;) {                              // cmb_objectqueue_put, src/cmb_objectqueue.c:262-314
                    if (st->ring_len < st->ring_cap) {
                        if (t.use_pq) {                 // cmb_priorityqueue_put with the putter's priority (:237-262)
                            if (st->pq.push(0u, p.stamp, p.prio, 0u, pid, 0) == 0u) {
                                st->status |= TRIAL_ERR_QUEUE_OVERFLOW;
                            }
                        }
                        else {
                            st->ring[(st->ring_head + st->ring_len) % st->ring_cap] = p.stamp;
                        }
                        st->ring_len++;
                        if (t.record) time_weighted_sample(t.hist, (double)st->ring_len, s.now);
                        s.signal(0u, st->ring_len > 0u);
                        t.c[0] += 1u;
                        break;
                    }
                    s.wait_begin(1u, pid);
                    p.pc = 2u;
                    return;
    case 2:
                    sig = s.wait_end(1u, pid, sig);
                    if (sig != (int32_t)SIG_SUCCESS) {
                        guarded_note(t, sig, 3);
                        break;
                    }
                }
            }
            else {
                for (;;) {                              // cmb_objectqueue_get, :203-260
                    if (st->ring_len > 0u) {
                        double stamp;
                        if (t.use_pq) {                 // cmb_priorityqueue_get (:189-212): the heap's first
                            (void)st->pq.pop();
                            stamp = st->pq.slot[0].d;
                        }
                        else {
                            stamp = st->ring[st->ring_head];
                            st->ring_head = (st->ring_head + 1u) % st->ring_cap;
                        }
                        st->ring_len--;
                        if (t.record) time_weighted_sample(t.hist, (double)st->ring_len, s.now);
                        s.signal(1u, st->ring_len < st->ring_cap);
                        t.c[1] += 1u;
                        t.sum_wait = __dadd_rn(t.sum_wait, __dsub_rn(s.now, stamp));
                        break;
                    }
                    s.wait_begin(0u, pid);
                    p.pc = 3u;
                    return;
    case 3:
                    sig = s.wait_end(0u, pid, sig);
                    if (sig != (int32_t)SIG_SUCCESS) {
                        guarded_note(t, sig, 4);
                        break;
                    }
                }
            }
        }
    }
}

template <bool TRACE>
__global__ void __launch_bounds__(GUARDED_BLOCK)
guarded_kernel(const GuardedArgs a)
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
    t.record = a.record;
    t.use_pq = a.use_pq;
    t.hist.start();
    if (t.record) {
        t.hist.sample(0.0, 0.0);                        // the empty queue at t = 0
    }

    st->pq.clear();
    st->fel.clear();
    st->guard[0].clear();
    st->guard[1].clear();
    st->guard[2].clear();
    st->tool_holder = NO_HOLDER;
    st->buf_cap = st->buf_level = 0u;
    st->ring_cap = (uint32_t)a.capacity;
    st->ring_head = 0u;
    st->ring_len = 0u;
    st->guard_seq = 0u;
    st->n_ew = 0u;
    st->tool_observer = 0u;
    st->status = TRIAL_OK;

    // run_guarded_trial in ref_driver.c: priorities are drawn, and START events
    // scheduled, in creation order; then the nuisance; then the end event
    for (uint32_t i = 0u; i <= GUARDED_WORKERS; i++) {
        GenProc &p = st->proc[i];
        p.pc = 0u;
        p.status = PROC_CREATED;
        p.n_awaits = 0u;
        p.n_waiters = 0u;
        p.hold_handle = 0u;
        p.guard_key = 0u;
        p.stamp = 0.0;
        p.holds_pool = p.holds_tool = p.held = p.req = p.rem = p.initially_held = 0u;
        p.kind = (i < GUARDED_PUTTERS) ? 0u : (i < GUARDED_WORKERS) ? 1u : 2u;
        p.prio = (i < GUARDED_WORKERS) ? (int32_t)s.rng.dice(-5, 5) : 0;
        s.schedule(ACT_START, i, 0, s.now, p.prio);     // cmb_process_start
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
        s.now = ev.d;                                   // src/cmb_event.c:239-241
        if (TRACE) {
            if (pops < a.trace_cap) {
                a.trace_key[trial * a.trace_cap + pops] = ev.key;
                a.trace_time[trial * a.trace_cap + pops] = s.now;
            }
        }
        pops++;
        const uint32_t pid = ev.subj;
        switch (ev.act) {
        case ACT_START:                                 // src/cmb_process.c:115-122
            st->proc[pid].status = PROC_RUNNING;
            st->proc[pid].pc = 0u;
            guarded_body(s, t, pid, ev.arg);
            break;
        case ACT_WAKE_TIME:                             // :292-308
            (void)s.await_remove(st->proc[pid], AWAIT_TIME, ev.key);
            guarded_body(s, t, pid, ev.arg);
            break;
        case ACT_WAKE_RESOURCE:                         // src/cmb_resourceguard.c:168-180
            if (st->proc[pid].status == PROC_RUNNING) {
                guarded_body(s, t, pid, ev.arg);
            }
            break;
        case ACT_WAKE_INTERRUPT:                        // src/cmb_process.c:628-643
            s.cancel_awaiteds(pid);
            guarded_body(s, t, pid, ev.arg);
            break;
        case ACT_USER:                                  // g_end_event: stop everybody
            for (uint32_t i = 0u; i <= GUARDED_WORKERS; i++) {
                s.stop(i);
            }
            break;
        }
    }

    t.c[6] = st->ring_len;
    if (t.record) {                                     // recording_stop + cmb_timeseries_summarize
        t.hist.sample((double)st->ring_len, s.now);
        t.c[6] = (uint64_t)__double_as_longlong(t.hist.acc.m1);
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
