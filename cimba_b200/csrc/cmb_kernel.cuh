// cmb_kernel.cuh - what turns a model written against cmb_device.cuh into a launchable trial kernel.
//
// A model is a struct with
//     void run_trial(cmb::Sim &sim, const cmb::TrialIn &in);     everything the reference's run_trial does BEFORE
//                                                                cmb_event_queue_execute(): initialise resources,
//                                                                create and start processes, schedule an end event
//     void process(cmb::Sim &sim, uint32_t me, uint32_t kind, int64_t sig);     the process bodies
//     void event(cmb::Sim &sim, uint32_t action, uint32_t subject, int64_t arg); its own events
//     bool demand(cmb::Sim &sim, uint32_t id, uint32_t pid, int32_t ctx);       its own wait predicates
//     void finish(cmb::Sim &sim, cmb::TrialOut &out);            what run_trial does AFTER the event list ran dry
// CMB_EXPORT_MODEL(Model) at the end of the model's .cu file emits the two C entry points the library loads
// (cimba_b200_model_load): the launcher and the workspace size.
//
// One trial per thread; a thread takes trials grid-stride, so any number of trials runs on a grid that fits the
// machine.  The same kernel is the REPAIR pass of the fixed-capacity fast kernels: with `only_flagged` set it re-runs
// just the trials whose status word carries one of those bits (queue / wait-list overflow) - the general engine's
// containers grow, so the reference's CMB_UNLIMITED queue holds there too - and clears the bits.
#pragma once

#include "../../include/cimba_b200.h"
#include "cmb_device.cuh"

namespace cimba_b200 {
namespace cmb {

struct TrialIn {
    double   arr_mean, srv_mean;
    uint64_t num_objects;
    int32_t  servers;
    uint32_t num_params;
    double   params[16];
    uint64_t trial;             // global trial index
};

struct TrialOut {
    uint64_t objects;
    double   sum_wait;
    uint32_t max_queue;
    uint64_t counters[8];
};

struct LaunchArgs {
    uint64_t master_seed, first_trial, num_trials, num_objects;
    int32_t  servers;
    uint32_t only_flagged;      // 0 = every trial; else only trials with (status & only_flagged) != 0
    const double *arr_mean, *srv_mean;
    uint64_t *events, *objects;
    double   *t_end, *sum_wait;
    uint32_t *status, *max_queue;
    uint64_t *counters;
    unsigned char *arena_base;  // workspace: [0, 256) = the allocation cursor, the rest = the arena
    unsigned long long arena_bytes;
    uint64_t  trace_cap;
    uint64_t *trace_key;
    double   *trace_time;
    unsigned long long *diag;
    uint32_t  num_params;
    double    params[16];
};

// one whole trial: the body of the reference's run_trial (benchmark/MM1_multi.c:91-125) around the dispatcher
template <class Model, bool TRACE>
CMB_FN void run_one_trial(Sim &sim, Model &m, const TrialIn &in, TrialOut &out,
                          uint64_t trace_cap, uint64_t *trace_key, double *trace_time)
{
    out.objects = 0u;
    out.sum_wait = 0.0;
    out.max_queue = 0u;
    for (int k = 0; k < 8; k++) out.counters[k] = 0u;
    m.run_trial(sim, in);
    execute<Model, TRACE>(sim, m, trace_cap, trace_key, trace_time);
    m.finish(sim, out);
}

#ifndef CMB_HOST_BUILD
constexpr int CMB_BLOCK = 64;

template <class Model, bool TRACE>
__global__ void __launch_bounds__(CMB_BLOCK)
trial_kernel(const LaunchArgs a)
{
    __shared__ ZigHot hot;
    stage_zig_hot(hot, true);
    __syncthreads();

    Arena arena;
    arena.base = a.arena_base + 256;
    arena.cursor = (unsigned long long *)a.arena_base;
    arena.bytes = a.arena_bytes;

    // The per-trial control block (cmb::Sim + the model) is the thread's stack frame, i.e. local memory: the hardware
    // interleaves it across the lanes word by word and - what decides it - caches it WRITE-BACK in L1.  Measured
    // (profiles/r02_engine.md): with the block in the HBM arena instead (compact per trial, but global stores write through
    // and the L1 hit rate fell from 98.6 % to 56 %) the long-scoreboard stall per issued instruction went from 3.4 to 108
    // and M/M/1 ran 2.1x slower.  Grown containers (heaps past their inline slots, the process table, queues) live in
    // the arena either way.
    Sim sim;
    Model m;

    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t trial = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; trial < a.num_trials; trial += stride) {
        if (a.only_flagged != 0u) {
            if ((a.status[trial] & a.only_flagged) == 0u) continue;
            if (a.diag != nullptr) atomicAdd(a.diag + 2, 1ull);
        }
        TrialIn in;
        TrialOut out;
        in.arr_mean = a.arr_mean[trial];
        in.srv_mean = a.srv_mean[trial];
        in.num_objects = a.num_objects;
        in.servers = a.servers;
        in.num_params = a.num_params;
        for (int k = 0; k < 16; k++) in.params[k] = a.params[k];
        in.trial = a.first_trial + trial;
        sim.init(fmix64(a.master_seed, a.first_trial + trial), &hot, arena);    // test/test_cimba.c:396
        run_one_trial<Model, TRACE>(sim, m, in, out, a.trace_cap,
                                    TRACE ? a.trace_key + trial * a.trace_cap : nullptr,
                                    TRACE ? a.trace_time + trial * a.trace_cap : nullptr);
        if (a.events)    a.events[trial] = sim.pops;
        if (a.objects)   a.objects[trial] = out.objects;
        if (a.t_end)     a.t_end[trial] = sim.now;
        if (a.sum_wait)  a.sum_wait[trial] = out.sum_wait;
        if (a.status)    a.status[trial] = sim.status;
        if (a.max_queue) a.max_queue[trial] = out.max_queue;
        if (a.counters) {
            for (int k = 0; k < 8; k++) a.counters[trial * 8u + k] = out.counters[k];
        }
    }
}
#endif  // CMB_HOST_BUILD

}  // namespace cmb
}  // namespace cimba_b200
