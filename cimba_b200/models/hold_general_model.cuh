// hold_general_model.cuh - the hold model (CIMBA_B200_MODEL_HOLD, oracle: ref_driver.c model 7) written against the
// authoring surface: `servers` worker processes in cmb_process_hold(exponential) loops, a ticker holding exactly 1.0,
// an end event that stops everybody.  The fused kernels (hold_deep.cuh: a warp-private 32-ary heap) are what runs this
// shape fast; this is the same model on the general engine - a thousand and more processes on the growable event
// list - as job->variant = CIMBA_B200_VARIANT_GENERAL, and what a model author would write.
#pragma once
#include "../csrc/cmb_kernel.cuh"

namespace cimba_b200 {
namespace models {

struct HoldGeneral {
    uint32_t workers;
    double   mean, sum_time;
    uint64_t wakeups, ticks;
    enum : uint32_t { WORKER, TICKER };
    enum : uint32_t { END_EVENT = cmb::ACT_CMB_USER };

    static uint64_t arena_bytes_per_trial(const cimba_b200_device_job &job)
    {
        const uint64_t n = (uint64_t)(job.servers < 8 ? 8 : job.servers) + 2u;
        return n * (2u * sizeof(cmb::Process) + 6u * sizeof(cmb::Tag) + 8u * sizeof(cmb::MapSlot) + 4u * sizeof(cmb::Node)) + 16384u;
    }

    CMB_FN void worker(cmb::Sim &sim, uint32_t me, int64_t sig)
    {
        HoldGeneral &m = *this;
        CMB_PROCESS_BEGIN
        for (;;) {
            CMB_PROCESS_HOLD_EXPONENTIAL(mean);
            wakeups += 1u;
            sum_time += cmb_time();
        }
        CMB_PROCESS_END
    }

    CMB_FN void ticker(cmb::Sim &sim, uint32_t me, int64_t sig)
    {
        HoldGeneral &m = *this;
        CMB_PROCESS_BEGIN
        for (;;) {
            CMB_PROCESS_HOLD(1.0);
            ticks += 1u;
        }
        CMB_PROCESS_END
    }

    CMB_FN void run_trial(cmb::Sim &sim, const cmb::TrialIn &in)
    {
        workers = (uint32_t)in.servers;
        mean = in.arr_mean;
        sum_time = 0.0;
        wakeups = ticks = 0u;
        (void)sim.process_reserve(workers + 1u);
        (void)sim.fel.reserve(sim.arena, workers + 3u);             // every process owns one pending event + the end event
        for (uint32_t i = 0u; i < workers; i++) cmb_process_start(cmb_process_create(WORKER, 0, i));
        cmb_process_start(cmb_process_create(TICKER, 0, 0u));
        (void)cmb_event_schedule(END_EVENT, cmb::NIL, 0, (double)in.num_objects, 0);
    }

    CMB_FN void process(cmb::Sim &sim, uint32_t me, uint32_t kind, int64_t sig)
    {
        if (kind == WORKER) worker(sim, me, sig);
        else ticker(sim, me, sig);
    }

    CMB_FN void event(cmb::Sim &sim, uint32_t action, uint32_t, int64_t)
    {
        HoldGeneral &m = *this;
        if (action == END_EVENT) {
            for (uint32_t i = 0u; i <= workers; i++) cmb_process_stop(i, 0);
        }
    }
    CMB_FN bool demand(cmb::Sim &, uint32_t, uint32_t, int32_t) { return false; }

    CMB_FN void finish(cmb::Sim &sim, cmb::TrialOut &out)
    {
        out.objects = wakeups;
        out.sum_wait = sum_time;
        out.counters[0] = wakeups;
        out.counters[1] = ticks;
        out.max_queue = sim.fel.cap();
    }
};

}  // namespace models
}  // namespace cimba_b200
