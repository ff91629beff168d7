// mm1_model.cuh - benchmark/MM1_multi.c:33-125 written against the device authoring surface (cmb_device.cuh).
// Compare with the reference file line by line: two process bodies, one run_trial.
// A template over the engine: MM1T<cmb::Sim> is the general-engine model, MM1T<cmb::StaticSim<2, 1>> the same text on the static
// tier (csrc/cmb_static.cuh: two processes, one queue, everything in registers and shared memory).
#pragma once
#include "../csrc/cmb_kernel.cuh"
#include "../csrc/cmb_static.cuh"

namespace cimba_b200 {
namespace models {

template <class S>
struct MM1T {
    typename S::queue_type queue;                             // struct simulation, MM1_multi.c:33-37
    uint32_t arrival, service;
    double   arr_mean, srv_mean;                        // struct trial, :39-45
    uint64_t num_objects, obj_cnt;
    double   sum_wait;
    uint64_t ui, stamp, object;                         // body locals that live across a blocking call
    enum : uint32_t { ARRIVAL, SERVICE };
    static CMB_FN constexpr uint32_t static_kind(uint32_t i) { return i == 0u ? ARRIVAL : SERVICE; }    // creation order (run_trial): what the static tier dispatches by
    static constexpr bool exponential_holds_only = true;     // no body draws: the static tier may look one variate ahead

    CMB_FN void arrivalfunc(S &sim, uint32_t me, int64_t sig)            // :52-68
    {
        MM1T &m = *this;
        CMB_PROCESS_BEGIN
        for (ui = 0u; ui < num_objects; ui++) {
            CMB_PROCESS_HOLD_EXPONENTIAL(arr_mean);
            stamp = (uint64_t)__double_as_longlong(cmb_time());
            CMB_OBJECTQUEUE_PUT(queue, stamp);
        }
        CMB_PROCESS_END
    }

    CMB_FN void servicefunc(S &sim, uint32_t me, int64_t sig)            // :70-89
    {
        MM1T &m = *this;
        CMB_PROCESS_BEGIN
        for (;;) {
            CMB_OBJECTQUEUE_GET(queue, object);
            CMB_PROCESS_HOLD_EXPONENTIAL(srv_mean);
            sum_wait += cmb_time() - __longlong_as_double((long long)object);
            obj_cnt += 1u;
        }
        CMB_PROCESS_END
    }

    CMB_FN void run_trial(S &sim, const cmb::TrialIn &in)               // :91-111
    {
        arr_mean = in.arr_mean;
        srv_mean = in.srv_mean;
        num_objects = in.num_objects;
        obj_cnt = 0u;
        sum_wait = 0.0;
        cmb_objectqueue_initialize(queue, CMB_UNLIMITED);
        arrival = cmb_process_create(ARRIVAL, 0, 0u);
        cmb_process_start(arrival);
        service = cmb_process_create(SERVICE, 0, 0u);
        cmb_process_start(service);
    }

    CMB_FN void process(S &sim, uint32_t me, uint32_t kind, int64_t sig)
    {
        if (kind == ARRIVAL) arrivalfunc(sim, me, sig);
        else servicefunc(sim, me, sig);
    }
    CMB_FN void event(S &, uint32_t, uint32_t, int64_t) {}
    CMB_FN bool demand(S &, uint32_t, uint32_t, int32_t) { return false; }

    CMB_FN void finish(S &sim, cmb::TrialOut &out)                       // :115-124
    {
        MM1T &m = *this;
        cmb_process_stop(service, 0);
        out.objects = obj_cnt;
        out.sum_wait = sum_wait;
    }
};

using MM1 = MM1T<cmb::Sim>;

}  // namespace models
}  // namespace cimba_b200
