// cheese_model.cuh - the reference's own pool test, test/test_resourcepool.c:50-305, as it stands, written against the
// authoring surface: three mice (cmb_process_priority_set + cmb_resourcepool_acquire of 1..10 units), two rats
// (cmb_resourcepool_preempt), a cat interrupting a random rodent (cmb_random_flip picks INTERRUPTED or a signal in 10..100),
// the pool's usage history on, an end event that stops everybody.  With 20 units, 100 time units and the reference's seed
// it reproduces test/reference/resourcepool.txt ("N 120  Mean 19.77  StdDev 1.147  ...") - the golden file that depends
// on the holders' tie-break by process address: the test creates its processes one after the other, so
// address order is creation order, which is what the index-based key (pid + 1) gives.
// Oracle: the same test against the reference's API, oracle/ref_build/ref_driver.c model 18.
#pragma once
#include "../csrc/cmb_kernel.cuh"

namespace cimba_b200 {
namespace models {

struct Cheese {
    cmb::resourcepool cheese;
    uint64_t successes;
    double   sum_time;
    enum : uint32_t { MOUSE, RAT, CAT };
    enum : uint32_t { END_EVENT = cmb::ACT_CMB_USER };
    static constexpr uint32_t MICE = 3u, RATS = 2u, RODENTS = 5u;

    // one rodent: u[0] = amount held (the body's local), u[1] = the amount of the call in progress
    CMB_FN void rodent(cmb::Sim &sim, uint32_t me, int64_t sig, bool rat)
    {
        Cheese &m = *this;
        CMB_PROCESS_BEGIN
        sim.proc[me].u[0] = 0u;
        for (;;) {
            sim.proc[me].u[1] = (uint64_t)cmb_random_dice(1, 10);
            if (rat) {
                CMB_RESOURCEPOOL_PREEMPT(cheese, sim.proc[me].u[1]);
            }
            else {
                cmb_process_priority_set(me, cmb_random_dice(-10, 10));
                CMB_RESOURCEPOOL_ACQUIRE(cheese, sim.proc[me].u[1]);
            }
            if (sig == CMB_PROCESS_SUCCESS) {
                sim.proc[me].u[0] += sim.proc[me].u[1];
                successes += 1u;
                sum_time += cmb_time();
                CMB_PROCESS_HOLD_EXPONENTIAL(1.0);
                if (sig == CMB_PROCESS_SUCCESS) {
                    uint64_t rel = (uint64_t)cmb_random_dice(1, 10);
                    if (rel > sim.proc[me].u[0]) rel = sim.proc[me].u[0];
                    CMB_RESOURCEPOOL_RELEASE(cheese, rel);
                    sim.proc[me].u[0] -= rel;
                }
                else if (sig == CMB_PROCESS_PREEMPTED) {
                    sim.proc[me].u[0] = 0u;
                }
            }
            else if (sig == CMB_PROCESS_PREEMPTED) {
                sim.proc[me].u[0] = 0u;
            }
            CMB_PROCESS_HOLD_EXPONENTIAL(1.0);
            if (sig == CMB_PROCESS_PREEMPTED) sim.proc[me].u[0] = 0u;
        }
        CMB_PROCESS_END
    }

    CMB_FN void cat(cmb::Sim &sim, uint32_t me, int64_t sig)
    {
        Cheese &m = *this;
        CMB_PROCESS_BEGIN
        for (;;) {
            CMB_PROCESS_HOLD_EXPONENTIAL(1.0);
            {
                const uint32_t victim = (uint32_t)cmb_random_dice(0, (long long)RODENTS - 1);
                const int64_t loud = cmb_random_dice(10, 100);
                cmb_process_interrupt(victim, cmb_random_flip() ? CMB_PROCESS_INTERRUPTED : loud, 0);
            }
        }
        CMB_PROCESS_END
    }

    CMB_FN void run_trial(cmb::Sim &sim, const cmb::TrialIn &in)                // test_pool, :307-379
    {
        successes = 0u;
        sum_time = 0.0;
        cmb_resourcepool_initialize(cheese, (uint64_t)in.servers);
        cmb_resourcepool_start_recording(cheese);
        for (uint32_t i = 0u; i <= RODENTS; i++) {
            const int64_t pri = cmb_random_dice(-5, 5);
            cmb_process_start(cmb_process_create(i < MICE ? MOUSE : (i < RODENTS ? RAT : CAT), pri, i));
        }
        (void)cmb_event_schedule(END_EVENT, cmb::NIL, 0, (double)in.num_objects, 0);
    }

    CMB_FN void process(cmb::Sim &sim, uint32_t me, uint32_t kind, int64_t sig)
    {
        if (kind == CAT) cat(sim, me, sig);
        else rodent(sim, me, sig, kind == RAT);
    }

    CMB_FN void event(cmb::Sim &sim, uint32_t action, uint32_t, int64_t)       // end_sim_evt, :50-72
    {
        Cheese &m = *this;
        if (action == END_EVENT) {
            for (uint32_t i = 0u; i <= RODENTS; i++) cmb_process_stop(i, 0);
        }
    }
    CMB_FN bool demand(cmb::Sim &, uint32_t, uint32_t, int32_t) { return false; }

    CMB_FN void finish(cmb::Sim &sim, cmb::TrialOut &out)
    {
        cmb_resourcepool_stop_recording(cheese);
        const WtdAcc &h = cheese.history.acc;           // what cmb_timeseries_summarize makes of the stored history
        out.counters[0] = h.count;
        out.counters[1] = (uint64_t)__double_as_longlong(h.min);
        out.counters[2] = (uint64_t)__double_as_longlong(h.max);
        out.counters[3] = (uint64_t)__double_as_longlong(h.m1);
        out.counters[4] = (uint64_t)__double_as_longlong(h.m2);
        out.counters[5] = (uint64_t)__double_as_longlong(h.m3);
        out.counters[6] = (uint64_t)__double_as_longlong(h.m4);
        out.counters[7] = (uint64_t)__double_as_longlong(h.wsum);
        out.objects = successes;
        out.sum_wait = sum_time;
        out.max_queue = (uint32_t)h.count;
    }
};

}  // namespace models
}  // namespace cimba_b200
