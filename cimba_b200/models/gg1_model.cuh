// gg1_model.cuh - BASELINE config 4 against the device authoring surface: benchmark/MM1_multi.c
// with Erlang-2 inter-arrival times and normal service times redrawn while negative (oracle: ref_driver.c model 1).
// A template over the engine, like mm1_model.cuh: GG1T<cmb::Sim> and GG1T<cmb::StaticSim<2, 1>>.
#pragma once
#include "../csrc/cmb_kernel.cuh"
#include "../csrc/cmb_static.cuh"

namespace cimba_b200 {
namespace models {

template <class S>
struct GG1T {
    typename S::queue_type queue;
    uint32_t arrival, service;
    double   arr_mean, srv_mean;
    uint64_t num_objects, obj_cnt;
    double   sum_wait;
    uint64_t ui, stamp, object;
    enum : uint32_t { ARRIVAL, SERVICE };
    static CMB_FN constexpr uint32_t static_kind(uint32_t i) { return i == 0u ? ARRIVAL : SERVICE; }    // creation order (run_trial): what the static tier dispatches by

    CMB_FN void arrivalfunc(S &sim, uint32_t me, int64_t sig)
    {
        GG1T &m = *this;
        CMB_PROCESS_BEGIN
        for (ui = 0u; ui < num_objects; ui++) {
            CMB_PROCESS_HOLD_SAMPLED(ARRIVAL);          // cmb_process_hold(<Erlang-2>): drawn by the dispatcher, see sample()
            stamp = (uint64_t)__double_as_longlong(cmb_time());
            CMB_OBJECTQUEUE_PUT(queue, stamp);
        }
        CMB_PROCESS_END
    }

    CMB_FN void servicefunc(S &sim, uint32_t me, int64_t sig)
    {
        GG1T &m = *this;
        CMB_PROCESS_BEGIN
        for (;;) {
            CMB_OBJECTQUEUE_GET(queue, object);
            CMB_PROCESS_HOLD_SAMPLED(SERVICE);          // cmb_process_hold(<normal, redrawn while negative>)
            sum_wait += cmb_time() - __longlong_as_double((long long)object);
            obj_cnt += 1u;
        }
        CMB_PROCESS_END
    }

    // the durations of the two holds, as SURVEY.md section 8d-4 defines them
    CMB_FN double sample(S &sim, uint32_t which)
    {
        if (which == ARRIVAL) return cmb_random_erlang(2u, 0.5 * arr_mean);
        double v;
        do {
            v = cmb_random_normal(srv_mean, 0.25 * srv_mean);
        } while (v < 0.0);
        return v;
    }

    CMB_FN void run_trial(S &sim, const cmb::TrialIn &in)
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

    CMB_FN void finish(S &, cmb::TrialOut &out)
    {
        out.objects = obj_cnt;
        out.sum_wait = sum_wait;
    }
};

using GG1 = GG1T<cmb::Sim>;

}  // namespace models
}  // namespace cimba_b200
