// examples/tandem_model.cuh - a model that exists NOWHERE in the library: two stations in tandem with a bounded buffer
// between them (a putter blocks on a full cmb_objectqueue), written from scratch against the authoring surface
// (cimba_b200/csrc/cmb_device.cuh).  examples/tandem_user_model.cu exports it as a loadable model library; the same
// model written against the reference's API is oracle/ref_build/ref_driver.c model 17.  A template over the engine: the same
// text runs on the general engine (cmb::Sim) and, three processes and two queues being all it has, on the static tier
// (cmb::StaticSim<3, 2>, cimba_b200/csrc/cmb_static.cuh).
#pragma once
#include "../cimba_b200/csrc/cmb_kernel.cuh"
#include "../cimba_b200/csrc/cmb_static.cuh"

namespace tandem_example {
using namespace cimba_b200;

template <class S>
struct TandemT {
    typename S::queue_type first, second;                     // unlimited in front of station 1, `servers` places in front of station 2
    double   arr_mean, srv_mean;
    uint64_t num_objects, done;
    double   sum_wait;
    uint64_t ui, stamp, at1, at2;
    enum : uint32_t { SOURCE, STATION1, STATION2 };
    static CMB_FN constexpr uint32_t static_kind(uint32_t i) { return i == 0u ? SOURCE : (i == 1u ? STATION1 : STATION2); }   // creation order, for the static tier's dispatcher

    CMB_FN void source(S &sim, uint32_t me, int64_t sig)
    {
        TandemT &m = *this;
        CMB_PROCESS_BEGIN
        for (ui = 0u; ui < num_objects; ui++) {
            CMB_PROCESS_HOLD_EXPONENTIAL(arr_mean);
            stamp = (uint64_t)__double_as_longlong(cmb_time());
            CMB_OBJECTQUEUE_PUT(first, stamp);
        }
        CMB_PROCESS_END
    }
    CMB_FN void station1(S &sim, uint32_t me, int64_t sig)
    {
        TandemT &m = *this;
        CMB_PROCESS_BEGIN
        for (;;) {
            CMB_OBJECTQUEUE_GET(first, at1);
            CMB_PROCESS_HOLD(cmb_random_exponential(srv_mean));
            CMB_OBJECTQUEUE_PUT(second, at1);           // blocks while the buffer is full
        }
        CMB_PROCESS_END
    }
    CMB_FN void station2(S &sim, uint32_t me, int64_t sig)
    {
        TandemT &m = *this;
        CMB_PROCESS_BEGIN
        for (;;) {
            CMB_OBJECTQUEUE_GET(second, at2);
            CMB_PROCESS_HOLD(cmb_random_uniform(0.5 * srv_mean, 1.5 * srv_mean));
            sum_wait += cmb_time() - __longlong_as_double((long long)at2);
            done += 1u;
        }
        CMB_PROCESS_END
    }

    CMB_FN void run_trial(S &sim, const cmb::TrialIn &in)
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
    CMB_FN void process(S &sim, uint32_t me, uint32_t kind, int64_t sig)
    {
        if (kind == SOURCE) source(sim, me, sig);
        else if (kind == STATION1) station1(sim, me, sig);
        else station2(sim, me, sig);
    }
    CMB_FN void event(S &, uint32_t, uint32_t, int64_t) {}
    CMB_FN bool demand(S &, uint32_t, uint32_t, int32_t) { return false; }
    CMB_FN void finish(S &, cmb::TrialOut &out)
    {
        out.objects = done;
        out.sum_wait = sum_wait;
    }
};
using Tandem = TandemT<cmb::Sim>;     // on the general engine; TandemT<cmb::StaticSim<3, 2>> is the static tier's (tandem_static_user_model.cu)
}  // namespace tandem_example
