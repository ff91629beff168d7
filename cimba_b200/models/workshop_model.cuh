// workshop_model.cuh - the reference's buffer and resource tests written against the authoring surface.
//
// Workshop<PLAIN>: test/test_buffer.c and test/test_resource.c in one world - fillers and drainers moving random amounts
// through a cmb_buffer of capacity `servers`, a polite and a pre-empting worker sharing one cmb_resource, a nuisance
// interrupting all six, an end event stopping everybody.  PLAIN = false: two fillers, two drainers, two workers, amounts
// 1..8 (model 5); PLAIN = true: test/test_buffer.c as it stands - three fillers, three drainers, amounts 1..15, the level
// history on (model 12, golden file test/reference/buffer.txt).
// Tool: test/test_resource.c as it stands - three targets with random priorities and a pre-empter (priority 0) on one
// cmb_resource with its history on (model 14, golden file test/reference/resource.txt).
// Oracle: oracle/ref_build/ref_driver.c run_buffer_trial / run_resource_trial (the counters are described there).
#pragma once
#include "../csrc/cmb_kernel.cuh"

namespace cimba_b200 {
namespace models {

template <bool PLAIN>
struct Workshop {
    cmb::buffer   store;
    cmb::resource tool;
    uint64_t counter[8];
    double   sum_wait, put_mean, get_mean;
    enum : uint32_t { FILLER, DRAINER, WORKER, NUISANCE };
    enum : uint32_t { END_EVENT = cmb::ACT_CMB_USER };
    static constexpr uint32_t PROCS = 6u;
    static constexpr long long AMOUNT_MAX = PLAIN ? 15 : 8;

    CMB_FN void note(int64_t sig)
    {
        if (sig != CMB_PROCESS_SUCCESS) counter[6] += (uint64_t)sig;
    }

    // u[0] = the amount offered / wanted, u[1] = what the call left of it
    CMB_FN void filler(cmb::Sim &sim, uint32_t me, int64_t sig)
    {
        Workshop &m = *this;
        CMB_PROCESS_BEGIN
        for (;;) {
            CMB_PROCESS_HOLD_EXPONENTIAL(put_mean);
            note(sig);
            sim.proc[me].u[0] = (uint64_t)cmb_random_dice(1, AMOUNT_MAX);
            sim.proc[me].u[1] = sim.proc[me].u[0];
            CMB_BUFFER_PUT(store, sim.proc[me].u[1]);
            counter[0] += sim.proc[me].u[0] - sim.proc[me].u[1];
            if (sig != CMB_PROCESS_SUCCESS) {
                counter[2] += 1u;
                note(sig);
            }
        }
        CMB_PROCESS_END
    }

    CMB_FN void drainer(cmb::Sim &sim, uint32_t me, int64_t sig)
    {
        Workshop &m = *this;
        CMB_PROCESS_BEGIN
        for (;;) {
            CMB_PROCESS_HOLD_EXPONENTIAL(get_mean);
            note(sig);
            sim.proc[me].u[1] = (uint64_t)cmb_random_dice(1, AMOUNT_MAX);
            CMB_BUFFER_GET(store, sim.proc[me].u[1]);
            counter[1] += sim.proc[me].u[1];
            if (sig != CMB_PROCESS_SUCCESS) {
                counter[3] += 1u;
                note(sig);
            }
        }
        CMB_PROCESS_END
    }

    // f[0] = when the tool was taken
    CMB_FN void worker(cmb::Sim &sim, uint32_t me, int64_t sig)
    {
        Workshop &m = *this;
        CMB_PROCESS_BEGIN
        for (;;) {
            if (me == 5u) CMB_RESOURCE_PREEMPT(tool);
            else CMB_RESOURCE_ACQUIRE(tool);
            if (sig == CMB_PROCESS_SUCCESS) {
                counter[4] += 1u;
                sim.proc[me].f[0] = cmb_time();
                CMB_PROCESS_HOLD_EXPONENTIAL(1.0);
                if (sig == CMB_PROCESS_PREEMPTED) {
                    counter[5] += 1u;
                    note(sig);
                }
                else {
                    note(sig);
                    CMB_RESOURCE_RELEASE(tool);
                    sum_wait = __dadd_rn(sum_wait, __dsub_rn(cmb_time(), sim.proc[me].f[0]));
                }
            }
            else {
                note(sig);
            }
            CMB_PROCESS_HOLD_EXPONENTIAL(1.0);
            note(sig);
        }
        CMB_PROCESS_END
    }

    CMB_FN void nuisance(cmb::Sim &sim, uint32_t me, int64_t sig)
    {
        Workshop &m = *this;
        CMB_PROCESS_BEGIN
        for (;;) {
            CMB_PROCESS_HOLD_EXPONENTIAL(1.0);
            {
                const uint32_t victim = (uint32_t)cmb_random_dice(0, (long long)PROCS - 1);
                const int64_t loud = cmb_random_dice(1, 10);
                const int64_t pri = cmb_random_dice(-5, 5);
                cmb_process_interrupt(victim, loud, pri);
            }
        }
        CMB_PROCESS_END
    }

    CMB_FN void run_trial(cmb::Sim &sim, const cmb::TrialIn &in)
    {
        for (uint32_t i = 0u; i < 8u; i++) counter[i] = 0u;
        sum_wait = 0.0;
        put_mean = in.arr_mean;
        get_mean = in.srv_mean;
        cmb_buffer_initialize(store, (uint64_t)in.servers);
        if (PLAIN) cmb_buffer_recording_start(store);
        cmb_resource_initialize(tool);
        const uint32_t fillers = PLAIN ? 3u : 2u, drainers = PLAIN ? 3u : 2u;
        for (uint32_t i = 0u; i < PROCS; i++) {
            const int64_t pri = cmb_random_dice(-5, 5);
            cmb_process_start(cmb_process_create(i < fillers ? FILLER : (i < fillers + drainers ? DRAINER : WORKER), pri, i));
        }
        cmb_process_start(cmb_process_create(NUISANCE, 0, PROCS));
        (void)cmb_event_schedule(END_EVENT, cmb::NIL, 0, (double)in.num_objects, 0);
    }

    CMB_FN void process(cmb::Sim &sim, uint32_t me, uint32_t kind, int64_t sig)
    {
        if (kind == FILLER) filler(sim, me, sig);
        else if (kind == DRAINER) drainer(sim, me, sig);
        else if (kind == WORKER) worker(sim, me, sig);
        else nuisance(sim, me, sig);
    }

    CMB_FN void event(cmb::Sim &sim, uint32_t action, uint32_t, int64_t)
    {
        Workshop &m = *this;
        if (action == END_EVENT) {
            for (uint32_t i = 0u; i <= PROCS; i++) cmb_process_stop(i, 0);
        }
    }
    CMB_FN bool demand(cmb::Sim &, uint32_t, uint32_t, int32_t) { return false; }

    CMB_FN void finish(cmb::Sim &sim, cmb::TrialOut &out)
    {
        counter[7] = cmb_buffer_level(store);
        out.max_queue = sim.fel_high;
        if (PLAIN) {
            cmb_buffer_recording_stop(store);
            counter[4] = (uint64_t)__double_as_longlong(store.history.acc.m1);
            out.max_queue = (uint32_t)store.history.acc.count;
        }
        for (uint32_t i = 0u; i < 8u; i++) out.counters[i] = counter[i];
        out.objects = counter[1];
        out.sum_wait = sum_wait;
    }
};

struct Tool {
    cmb::resource res;
    uint64_t counter[8];
    double   sum_wait;
    enum : uint32_t { TARGET, PREEMPTER };
    enum : uint32_t { END_EVENT = cmb::ACT_CMB_USER };

    CMB_FN void target(cmb::Sim &sim, uint32_t me, int64_t sig)
    {
        Tool &m = *this;
        CMB_PROCESS_BEGIN
        for (;;) {
            CMB_RESOURCE_ACQUIRE(res);
            if (sig == CMB_PROCESS_SUCCESS) {
                counter[0] += 1u;
                sim.proc[me].f[0] = cmb_time();
                CMB_PROCESS_HOLD_EXPONENTIAL(1.0);
                if (sig == CMB_PROCESS_SUCCESS) {
                    CMB_RESOURCE_RELEASE(res);
                    sum_wait = __dadd_rn(sum_wait, __dsub_rn(cmb_time(), sim.proc[me].f[0]));
                }
                else {
                    counter[1] += 1u;
                    if (counter[5] == 0u) {
                        counter[4] = (uint64_t)__double_as_longlong(cmb_time());
                        counter[5] = (uint64_t)me + 1u;
                    }
                }
            }
            CMB_PROCESS_HOLD_EXPONENTIAL(1.0);
        }
        CMB_PROCESS_END
    }

    CMB_FN void preempter(cmb::Sim &sim, uint32_t me, int64_t sig)
    {
        Tool &m = *this;
        CMB_PROCESS_BEGIN
        for (;;) {
            CMB_RESOURCE_PREEMPT(res);
            counter[2] += 1u;
            CMB_PROCESS_HOLD_EXPONENTIAL(1.0);
            CMB_RESOURCE_RELEASE(res);
            CMB_PROCESS_HOLD_EXPONENTIAL(1.0);
        }
        CMB_PROCESS_END
    }

    CMB_FN void run_trial(cmb::Sim &sim, const cmb::TrialIn &in)
    {
        for (uint32_t i = 0u; i < 8u; i++) counter[i] = 0u;
        sum_wait = 0.0;
        cmb_resource_initialize(res);
        cmb_resource_start_recording(res);
        for (uint32_t i = 0u; i < 3u; i++) {
            const int64_t pri = cmb_random_dice(-5, 5);
            cmb_process_start(cmb_process_create(TARGET, pri, i));
        }
        cmb_process_start(cmb_process_create(PREEMPTER, 0, 3u));
        (void)cmb_event_schedule(END_EVENT, cmb::NIL, 0, (double)in.num_objects, 0);
    }

    CMB_FN void process(cmb::Sim &sim, uint32_t me, uint32_t kind, int64_t sig)
    {
        if (kind == TARGET) target(sim, me, sig);
        else preempter(sim, me, sig);
    }

    CMB_FN void event(cmb::Sim &sim, uint32_t action, uint32_t, int64_t)
    {
        Tool &m = *this;
        if (action == END_EVENT) {
            for (uint32_t i = 0u; i < 4u; i++) cmb_process_stop(i, 0);
        }
    }
    CMB_FN bool demand(cmb::Sim &, uint32_t, uint32_t, int32_t) { return false; }

    CMB_FN void finish(cmb::Sim &sim, cmb::TrialOut &out)
    {
        cmb_resource_stop_recording(res);
        counter[3] = (uint64_t)__double_as_longlong(res.history.acc.m1);
        for (uint32_t i = 0u; i < 8u; i++) out.counters[i] = counter[i];
        out.max_queue = (uint32_t)res.history.acc.count;
        out.objects = counter[0] + counter[2];
        out.sum_wait = sum_wait;
    }
};

}  // namespace models
}  // namespace cimba_b200
