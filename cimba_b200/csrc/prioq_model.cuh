// prioq_model.cuh - cmb_priorityqueue + cmb_condition (model 6).
//
// Workload: oracle/ref_build/ref_driver.c model 6, in the manner of the reference's
// test/test_priorityqueue.c and test/test_condition.c: two producers and a consumer on
// a bounded priority queue, a shuffler that looks items up / re-ranks / withdraws them
// by handle, a tide process that sets a level and signals a condition that two waiters
// watch with different thresholds, a nuisance interrupting all seven, an end event.
//
// Parity vehicle for the rest of SURVEY.md section 8a row a16:
//   cmb_priorityqueue_put/get        src/cmb_priorityqueue.c:189-284 (store ordered by
//                                    priority desc then handle asc, :43-54)
//   ..._position/cancel/reprioritize :286-320, include/cmb_priorityqueue.h:152-185
//   cmb_condition_wait/signal        src/cmb_condition.c:63-167: signal walks the guard
//       in ARRAY order, schedules a wake-up for every waiter whose predicate holds, then
//       removes them; its wake-up event strips the waiter's RESOURCE awaitable itself
#pragma once

#include "general.cuh"
#include "guarded_model.cuh"

namespace cimba_b200 {

constexpr uint32_t PRIOQ_PROCS = 7u;   // 0,1 producers; 2 consumer; 3 shuffler; 4 tide; 5,6 waiters; 7 nuisance

__device__ __forceinline__ void prioq_note(GuardedTally &t, int32_t sig)
{
    if (sig != (int32_t)SIG_SUCCESS) {
        t.c[6] += (uint64_t)(int64_t)sig;
    }
}

__device__ __forceinline__ int32_t prioq_threshold(uint32_t pid)
{
    return pid == 5u ? 2 : 4;
}

// cmb_condition_signal, src/cmb_condition.c:120-167
__device__ uint32_t prioq_condition_signal(GeneralSim &s)
{
    GeneralState *st = s.st;
    GuardHeap &cv = st->guard[2];
    uint32_t hit[GEN_GUARD_CAP];
    uint32_t cnt = 0u;
    for (uint32_t k = 1u; k <= cv.count; k++) {
        const uint32_t pid = cv.slot[k].subj;
        if (st->level >= prioq_threshold(pid)) {
            hit[cnt++] = cv.slot[k].key;
            s.schedule(ACT_WAKE_CONDITION, pid, (int32_t)SIG_SUCCESS, s.now, st->proc[pid].prio);
        }
    }
    for (uint32_t k = 0u; k < cnt; k++) {
        cv.remove(hit[k]);
    }
    return cnt;
}

// cmb_priorityqueue_position, src/cmb_priorityqueue.c:286-320
__device__ uint32_t prioq_position(const PrioHeap &pq, uint32_t handle)
{
    const uint32_t at = pq.find(handle);
    if (at == 0u) {
        return 0u;
    }
    uint32_t ahead = 0u;
    for (uint32_t k = 1u; k <= pq.count; k++) {
        if (k != at && PrioOrder::before(pq.slot[k], pq.slot[at])) {
            ahead++;
        }
    }
    return ahead + 1u;
}

__device__ void prioq_body(GeneralSim &s, GuardedTally &t, uint32_t pid, int32_t sig)
{
    GeneralState *st = s.st;
    GenProc &p = st->proc[pid];
    switch (p.pc) {
    case 0:
        if (pid == PRIOQ_PROCS) {                       // nuisance
            for (;;) {
                s.hold_begin(pid, gp_exponential(s.rng, *s.hot, 1.0));
                p.pc = 10u;
                return;
    case 10:
                (void)s.hold_end(pid, sig);
                {
                    const uint32_t victim = (uint32_t)s.rng.dice(0, PRIOQ_PROCS - 1u);
                    const int32_t isig = (int32_t)s.rng.dice(1, 10);
                    const int32_t ipri = (int32_t)s.rng.dice(-5, 5);
                    s.interrupt(victim, isig, ipri);
                }
            }
        }
        if (pid < 2u) {                                 // producer
            for (;;) {
                s.hold_begin(pid, gp_exponential(s.rng, *s.hot, t.put_mean));
                p.pc = 20u;
                return;
    case 20:
                prioq_note(t, s.hold_end(pid, sig));
                p.req = (uint32_t)s.rng.dice(1, 9);     // the object: a weight
                p.rem = (uint32_t)(int32_t)s.rng.dice(-3, 3);   // its priority
                for (;;) {                              // cmb_priorityqueue_put
                    if (st->pq.count < st->pq_cap) {
                        p.held = st->pq.push(0u, 0.0, (int32_t)p.rem, 0u, 0u, (int32_t)p.req);
                        s.signal(0u, st->pq.count > 0u);
                        sig = (int32_t)SIG_SUCCESS;
                        break;
                    }
                    s.wait_begin(1u, pid);
                    p.pc = 21u;
                    return;
    case 21:
                    sig = s.wait_end(1u, pid, sig);
                    if (sig != (int32_t)SIG_SUCCESS) {
                        break;
                    }
                }
                if (sig == (int32_t)SIG_SUCCESS) {
                    t.c[0] += 1u;
                    st->last_handle[pid] = p.held;
                }
                else {
                    t.c[2] += 1u;
                    prioq_note(t, sig);
                }
            }
        }
        if (pid == 2u) {                                // consumer
            for (;;) {
                s.hold_begin(pid, gp_exponential(s.rng, *s.hot, t.get_mean));
                p.pc = 30u;
                return;
    case 30:
                prioq_note(t, s.hold_end(pid, sig));
                for (;;) {                              // cmb_priorityqueue_get
                    if (st->pq.count > 0u) {
                        st->pq.pop();
                        p.req = (uint32_t)st->pq.slot[0].arg;
                        s.signal(1u, st->pq.count < st->pq_cap);
                        sig = (int32_t)SIG_SUCCESS;
                        break;
                    }
                    s.wait_begin(0u, pid);
                    p.pc = 31u;
                    return;
    case 31:
                    sig = s.wait_end(0u, pid, sig);
                    if (sig != (int32_t)SIG_SUCCESS) {
                        break;
                    }
                }
                if (sig == (int32_t)SIG_SUCCESS) {
                    t.c[1] += p.req;
                    t.sum_wait = __dadd_rn(t.sum_wait, __dmul_rn(s.now, (double)p.req));
                }
                else {
                    t.c[2] += 1u;
                    prioq_note(t, sig);
                }
            }
        }
        if (pid == 3u) {                                // shuffler
            for (;;) {
                s.hold_begin(pid, gp_exponential(s.rng, *s.hot, 1.5));
                p.pc = 40u;
                return;
    case 40:
                prioq_note(t, s.hold_end(pid, sig));
                {
                    const uint32_t handle = st->last_handle[s.rng.dice(0, 1)];
                    if (handle == 0u) {
                        continue;
                    }
                    const uint32_t pos = prioq_position(st->pq, handle);
                    t.c[3] += pos;
                    if (pos > 0u) {
                        if (s.rng.dice(0, 1) == 1) {
                            st->pq.reprioritize(handle, 0.0, (int32_t)s.rng.dice(-3, 3));
                        }
                        else {
                            (void)st->pq.remove(handle);
                            t.c[3] += 1000u;
                        }
                    }
                }
            }
        }
        if (pid == 4u) {                                // tide
            for (;;) {
                s.hold_begin(pid, gp_exponential(s.rng, *s.hot, 1.0));
                p.pc = 50u;
                return;
    case 50:
                prioq_note(t, s.hold_end(pid, sig));
                st->level = (int32_t)s.rng.dice(0, 5);
                t.c[4] += prioq_condition_signal(s);
            }
        }
        for (;;) {                                      // waiters
            {
                bool through;
                through = true;
                while (st->level < prioq_threshold(pid)) {
                    s.wait_begin(2u, pid);              // cmb_condition_wait = cmb_resourceguard_wait
                    p.pc = 60u;
                    return;
    case 60:
                    through = true;
                    sig = s.wait_end(2u, pid, sig);
                    if (sig != (int32_t)SIG_SUCCESS) {
                        prioq_note(t, sig);
                        through = false;
                        break;
                    }
                }
                if (through) {
                    t.c[5] += 1u;
                }
            }
            s.hold_begin(pid, gp_exponential(s.rng, *s.hot, 1.0));
            p.pc = 61u;
            return;
    case 61:
            prioq_note(t, s.hold_end(pid, sig));
        }
    }
}

template <bool TRACE>
__global__ void __launch_bounds__(GUARDED_BLOCK)
prioq_kernel(const GuardedArgs a)
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

    st->fel.clear();
    st->guard[0].clear();
    st->guard[1].clear();
    st->guard[2].clear();
    st->holders.clear();
    st->pq.clear();
    st->pq_cap = (uint32_t)a.capacity;
    st->last_handle[0] = st->last_handle[1] = 0u;
    st->level = 0;
    st->pool_cap = st->pool_in_use = 0u;
    st->buf_cap = st->buf_level = 0u;
    st->tool_holder = NO_HOLDER;
    st->guard_seq = 0u;
    st->n_ew = 0u;
    st->tool_observer = 0u;
    st->status = TRIAL_OK;
    st->ring_cap = 1u;
    st->ring_head = st->ring_len = 0u;

    for (uint32_t i = 0u; i <= PRIOQ_PROCS; i++) {
        GenProc &p = st->proc[i];
        p.pc = 0u;
        p.status = PROC_CREATED;
        p.kind = i;
        p.n_awaits = 0u;
        p.n_waiters = 0u;
        p.hold_handle = p.guard_key = 0u;
        p.stamp = 0.0;
        p.holds_pool = p.holds_tool = p.held = p.req = p.rem = p.initially_held = 0u;
        p.prio = (i < PRIOQ_PROCS) ? (int32_t)s.rng.dice(-5, 5) : 0;
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
        case ACT_WAKE_CONDITION:
            (void)s.await_remove_any(st->proc[pid], AWAIT_RESOURCE);
            run = st->proc[pid].status == PROC_RUNNING;
            break;
        case ACT_WAKE_INTERRUPT:
            s.cancel_awaiteds(pid);
            run = true;
            break;
        case ACT_USER:
            for (uint32_t i = 0u; i <= PRIOQ_PROCS; i++) {
                s.stop(i);
            }
            break;
        }
        if (run) {
            prioq_body(s, t, pid, ev.arg);
        }
    }

    t.c[7] = st->pq.count;
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
