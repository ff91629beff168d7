// examples/tandem_model.cuh - a model that exists NOWHERE in the library: two stations in tandem with a bounded buffer
// between them (a putter blocks on a full cmb_objectqueue), written from scratch against the authoring surface
// (cimba_b200/csrc/cmb_device.cuh).  examples/tandem_user_model.cu exports it as a loadable model library; the same
// model written against the reference's API is oracle/ref_build/ref_driver.c model 17.
#pragma once
#include "../cimba_b200/csrc/cmb_kernel.cuh"

namespace tandem_example {
using namespace cimba_b200;

struct Tandem {
    cmb::objectqueue first, second;                     // unlimited in front of station 1, `servers` places in front of station 2
    double   arr_mean, srv_mean;
    uint64_t num_objects, done;
    double   sum_wait;
    uint64_t ui, stamp, at1, at2;
    enum : uint32_t { SOURCE, STATION1, STATION2 };

    CMB_FN void source(cmb::Sim &sim, uint32_t me, int64_t sig)
    {
        Tandem &m = *this;
        CMB_PROCESS_BEGIN
        for (ui = 0u; ui < num_objects; ui++) {
            CMB_PROCESS_HOLD_EXPONENTIAL(arr_mean);
            stamp = (uint64_t)__double_as_longlong(cmb_time());
            CMB_OBJECTQUEUE_PUT(first, stamp);
        }
        CMB_PROCESS_END
    }
    CMB_FN void station1(cmb::Sim &sim, uint32_t me, int64_t sig)
    {
        Tandem &m = *this;
        CMB_PROCESS_BEGIN
        for (;;) {
            CMB_OBJECTQUEUE_GET(first, at1);
            CMB_PROCESS_HOLD(cmb_random_exponential(srv_mean));
            CMB_OBJECTQUEUE_PUT(second, at1);           // blocks while the buffer is full
        }
        CMB_PROCESS_END
    }
    CMB_FN void station2(cmb::Sim &sim, uint32_t me, int64_t sig)
    {
        Tandem &m = *this;
        CMB_PROCESS_BEGIN
        for (;;) {
            CMB_OBJECTQUEUE_GET(second, at2);
            CMB_PROCESS_HOLD(cmb_random_uniform(0.5 * srv_mean, 1.5 * srv_mean));
            sum_wait += cmb_time() - __longlong_as_double((long long)at2);
            done += 1u;
        }
        CMB_PROCESS_END
    }

    CMB_FN void run_trial(cmb::Sim &sim, const cmb::TrialIn &in)
    {
        arr_mean = in.arr_mean;
        srv_mean = in.srv_mean;
        num_objects = in.num_objects;
        done = 0u;
        sum_wait = 0.0;
        cmb_objectqueue_initialize(first, CMB_UNLIMITED);
        cmb_objectqueue_initialize(second, (uint64_t)in.servers);
        cmb_process_start(cmb_process_create(SOURCE, 0, 0u));
        cmb_process_start(cmb_process_create(STATION1, 0, 0u));
        cmb_process_start(cmb_process_create(STATION2, 0, 0u));
    }
    CMB_FN void process(cmb::Sim &sim, uint32_t me, uint32_t kind, int64_t sig)
    {
        if (kind == SOURCE) source(sim, me, sig);
        else if (kind == STATION1) station1(sim, me, sig);
        else station2(sim, me, sig);
    }
    CMB_FN void event(cmb::Sim &, uint32_t, uint32_t, int64_t) {}
    CMB_FN bool demand(cmb::Sim &, uint32_t, uint32_t, int32_t) { return false; }
    CMB_FN void finish(cmb::Sim &, cmb::TrialOut &out)
    {
        out.objects = done;
        out.sum_wait = sum_wait;
    }
};
}  // namespace tandem_example