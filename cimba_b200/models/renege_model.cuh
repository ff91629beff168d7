// renege_model.cuh - a service desk with impatient customers, in the manner of tutorial/tut_3_1.c (timers around a
// wait, reneging): `servers` customer PROCESSES (a thousand and more) think, then queue for one of servers / 8 clerks
// (a cmb_resourcepool) with a patience timer running; whoever is served in time clears the timer (an event cancelled
// by handle), whoever is not leaves the wait list (a wait-list entry removed by key).  An end event stops everybody.
// It is the general engine's stress test: the event list holds one or two thousand entries, handles are looked up
// through the hash map on every service, and the stop cascade cancels by pattern.  The oracle is the same model
// written against the reference's API: oracle/ref_build/ref_driver.c run_renege_trial.
#pragma once
#include "../csrc/cmb_kernel.cuh"

namespace cimba_b200 {
namespace models {

struct Renege {
    cmb::resourcepool clerks;
    uint32_t customers;
    double   think_mean, service_mean, patience_mean;
    uint64_t served, reneged, interrupted;
    double   sum_wait;
    enum : uint32_t { CUSTOMER };
    enum : uint32_t { END_OF_DAY = cmb::ACT_CMB_USER };
    static constexpr int64_t TIMER_RENEGING = 17;

    // growth memory of one trial: the process table, an event list of ~2 entries per customer with its key map, the
    // wait list, and as much again for the blocks the doubling leaves behind
    static uint64_t arena_bytes_per_trial(const cimba_b200_device_job &job)
    {
        const uint64_t n = (uint64_t)(job.servers < 8 ? 8 : job.servers);
        return n * (2u * sizeof(cmb::Process) + 8u * sizeof(cmb::Tag) + 16u * sizeof(cmb::MapSlot) + 8u * sizeof(cmb::Node)) + 16384u;
    }

    CMB_FN void customer(cmb::Sim &sim, uint32_t me, int64_t sig)
    {
        Renege &m = *this;
        CMB_PROCESS_BEGIN
        for (;;) {
            CMB_PROCESS_HOLD_EXPONENTIAL(think_mean);
            sim.proc[me].f[0] = cmb_time();                         // joined the line
            (void)cmb_process_timer_add(cmb_random_exponential(patience_mean), TIMER_RENEGING);
            CMB_RESOURCEPOOL_ACQUIRE(clerks, 1u);
            if (sig == CMB_PROCESS_SUCCESS) {
                cmb_process_timers_clear(me);
                sum_wait += cmb_time() - sim.proc[me].f[0];
                CMB_PROCESS_HOLD_EXPONENTIAL(service_mean);
                CMB_RESOURCEPOOL_RELEASE(clerks, 1u);
                served += 1u;
            }
            else if (sig == TIMER_RENEGING) {
                reneged += 1u;
            }
            else {
                interrupted += 1u;
            }
        }
        CMB_PROCESS_END
    }

    CMB_FN void run_trial(cmb::Sim &sim, const cmb::TrialIn &in)
    {
        customers = (uint32_t)in.servers;
        think_mean = in.arr_mean;
        service_mean = in.srv_mean;
        patience_mean = in.num_params > 0u ? in.params[0] : in.srv_mean;
        served = reneged = interrupted = 0u;
        sum_wait = 0.0;
        cmb_resourcepool_initialize(clerks, (uint64_t)((customers + 7u) / 8u));
        (void)sim.process_reserve(customers);
        for (uint32_t i = 0u; i < customers; i++) {
            const int64_t prio = cmb_random_dice(0, 3);
            const uint32_t pid = cmb_process_create(CUSTOMER, prio, i);
            cmb_process_start(pid);
        }
        (void)cmb_event_schedule(END_OF_DAY, cmb::NIL, 0, (double)in.num_objects, 0);
    }

    CMB_FN void process(cmb::Sim &sim, uint32_t me, uint32_t, int64_t sig) { customer(sim, me, sig); }

    CMB_FN void event(cmb::Sim &sim, uint32_t action, uint32_t, int64_t)
    {
        Renege &m = *this;
        if (action == END_OF_DAY) {
            for (uint32_t i = 0u; i < customers; i++) cmb_process_stop(i, 0);
        }
    }
    CMB_FN bool demand(cmb::Sim &, uint32_t, uint32_t, int32_t) { return false; }

    CMB_FN void finish(cmb::Sim &sim, cmb::TrialOut &out)
    {
        out.objects = served;
        out.sum_wait = sum_wait;
        out.counters[0] = served;
        out.counters[1] = reneged;
        out.counters[2] = interrupted;
        out.counters[3] = clerks.in_use;
        out.counters[4] = clerks.guard.heap.count;      // stale wait-list entries left by the stop cascade (quirk 2)
        out.counters[5] = sim.fel.exp;                  // how far the event list grew (log2 of its capacity)
        out.counters[6] = sim.fel.map_on;
        out.counters[7] = sim.nproc;
    }
};

}  // namespace models
}  // namespace cimba_b200
