// cmb_launch.cuh - host side of a model built on cmb_device.cuh: fill the kernel's arguments from the C-ABI job,
// size and reset the growth arena, launch.  Used by the library itself (its built-in general-engine models and the
// repair pass of the fast kernels) and, through CMB_EXPORT_MODEL, by a model compiled into a library of its own.
#pragma once

#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>

#include "../../include/cimba_b200.h"
#include "cmb_kernel.cuh"
#include "cmb_static.cuh"

namespace cimba_b200 {
namespace cmb {

// Growth memory of a launch.  A model may say what a trial needs (static uint64_t arena_bytes_per_trial(const
// cimba_b200_device_job &)); the default suits models with a handful of processes.
template <class Model, class = void>
struct ArenaNeed {
    static uint64_t per_trial(const cimba_b200_device_job &) { return 32768u; }
};
template <class Model>
struct ArenaNeed<Model, decltype((void)Model::arena_bytes_per_trial(*(const cimba_b200_device_job *)nullptr))> {
    static uint64_t per_trial(const cimba_b200_device_job &job) { return Model::arena_bytes_per_trial(job); }
};

constexpr uint64_t ARENA_HEADER = 256u;         // the allocation cursor lives in front of the arena

template <class Model>
uint64_t workspace_bytes_for(const cimba_b200_device_job &job)
{
    return ARENA_HEADER + job.num_trials * ArenaNeed<Model>::per_trial(job) + (64ull << 20);
}

// Launch Model over the job's trials on `stream`, growth arena = [arena, arena + arena_bytes) (device memory, 256-byte
// aligned; the first ARENA_HEADER bytes hold the cursor).  only_flagged = 0: every trial; else the repair pass.
// Returns a cudaError_t as int (0 = launched).
template <class Model>
int launch_model(const cimba_b200_device_job &job, unsigned char *arena, uint64_t arena_bytes, uint32_t only_flagged,
                 cudaStream_t stream)
{
    if (arena == nullptr || arena_bytes <= ARENA_HEADER) return (int)cudaErrorInvalidValue;
    LaunchArgs a{};
    a.master_seed = job.master_seed;
    a.first_trial = job.first_trial;
    a.num_trials = job.num_trials;
    a.num_objects = job.num_objects;
    a.servers = job.servers;
    a.only_flagged = only_flagged;
    a.arr_mean = job.arr_mean;
    a.srv_mean = job.srv_mean;
    a.events = job.events;
    a.objects = job.objects;
    a.t_end = job.t_end;
    a.sum_wait = job.sum_wait;
    a.status = job.status;
    a.max_queue = job.max_queue;
    a.counters = job.counters;
    a.arena_base = arena;
    a.arena_bytes = arena_bytes - ARENA_HEADER;
    a.trace_cap = job.trace_cap;
    a.trace_key = job.trace_key;
    a.trace_time = job.trace_time;
    a.diag = (unsigned long long *)job.diag;
    a.num_params = job.params != nullptr ? (job.num_params < 16u ? job.num_params : 16u) : 0u;
    for (uint32_t k = 0; k < a.num_params; k++) a.params[k] = job.params[k];

    cudaError_t e = cudaMemsetAsync(arena, 0, ARENA_HEADER, stream);
    if (e != cudaSuccess) return (int)e;
    const bool trace = job.trace_cap > 0u;
    const void *fn = trace ? (const void *)trial_kernel<Model, true> : (const void *)trial_kernel<Model, false>;
    // the per-trial control block (cmb::Sim + the model) is the kernel's stack frame: make room for it
    cudaFuncAttributes attr{};
    e = cudaFuncGetAttributes(&attr, fn);
    if (e != cudaSuccess) return (int)e;
    size_t limit = 0;
    (void)cudaDeviceGetLimit(&limit, cudaLimitStackSize);
    if (attr.localSizeBytes + 1024u > limit) {
        e = cudaDeviceSetLimit(cudaLimitStackSize, attr.localSizeBytes + 1024u);
        if (e != cudaSuccess) return (int)e;
    }
    int dev = 0, sms = 148;
    (void)cudaGetDevice(&dev);
    (void)cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint64_t wanted = (job.num_trials + CMB_BLOCK - 1) / CMB_BLOCK;
    const uint64_t resident = (uint64_t)sms * 16u;      // 16 CTAs of 64 threads per SM: trials beyond that are taken grid-stride
    const unsigned blocks = (unsigned)(wanted < resident ? wanted : resident);
    void *kargs[] = { (void *)&a };
    e = cudaLaunchKernel(fn, dim3(blocks), dim3(CMB_BLOCK), kargs, 0, stream);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaGetLastError();
}

// ---- the static tier (cmb_static.cuh): ModelT<StaticSim<NPROC, NQUEUE>> first, ModelT<Sim> for what it flags
constexpr uint32_t STATIC_REPAIR_BITS = CIMBA_B200_TRIAL_QUEUE_OVERFLOW | CIMBA_B200_TRIAL_FEL_OVERFLOW |
                                        CIMBA_B200_TRIAL_GUARD_OVERFLOW | CIMBA_B200_TRIAL_PROC_OVERFLOW;

inline uint32_t static_spill_cap(const cimba_b200_device_job &job)     // HBM ring entries per queue and trial (a power of two)
{
    const uint64_t c = job.queue_spill_cap;
    if (c == 0u) return 512u;
    return (c & (c - 1u)) == 0u && c <= (1ull << 26) ? (uint32_t)c : 0u;
}

inline uint64_t static_rings_bytes(const cimba_b200_device_job &job, int nqueue)
{
    const uint64_t b = job.num_trials * (uint64_t)nqueue * static_spill_cap(job) * sizeof(double);
    return (b + 255u) & ~(uint64_t)255u;
}

template <template <class> class ModelT, int NPROC, int NQUEUE, int NEVENT = 0>
uint64_t workspace_bytes_static(const cimba_b200_device_job &job)
{
    // the rings of the static kernel, then the growth arena of its repair pass (a share of the trials, not all of them)
    const uint64_t arena = ARENA_HEADER + (job.num_trials / 8u + 256u) * ArenaNeed<ModelT<Sim>>::per_trial(job) + (64ull << 20);
    return static_rings_bytes(job, NQUEUE) + arena;
}

template <template <class> class ModelT, int NPROC, int NQUEUE, int NEVENT = 0>
int launch_static_model(const cimba_b200_device_job &job, cudaStream_t stream)
{
    const uint32_t cap = static_spill_cap(job);
    const uint64_t rings = static_rings_bytes(job, NQUEUE);
    if (cap == 0u || job.workspace == nullptr || job.workspace_bytes < workspace_bytes_static<ModelT, NPROC, NQUEUE, NEVENT>(job))
        return (int)cudaErrorInvalidValue;
    StaticArgs sa{};
    LaunchArgs &a = sa.base;
    a.master_seed = job.master_seed;
    a.first_trial = job.first_trial;
    a.num_trials = job.num_trials;
    a.num_objects = job.num_objects;
    a.servers = job.servers;
    a.arr_mean = job.arr_mean;
    a.srv_mean = job.srv_mean;
    a.events = job.events;
    a.objects = job.objects;
    a.t_end = job.t_end;
    a.sum_wait = job.sum_wait;
    a.status = job.status;
    a.max_queue = job.max_queue;
    a.counters = job.counters;
    a.trace_cap = job.trace_cap;
    a.trace_key = job.trace_key;
    a.trace_time = job.trace_time;
    a.diag = (unsigned long long *)job.diag;
    a.num_params = job.params != nullptr ? (job.num_params < 16u ? job.num_params : 16u) : 0u;
    for (uint32_t k = 0; k < a.num_params; k++) a.params[k] = job.params[k];
    sa.spill = (double *)job.workspace;
    sa.spill_cap = cap;
    const uint64_t blocks = (job.num_trials + STATIC_BLOCK - 1) / STATIC_BLOCK;
    if (blocks == 0u || blocks > 0x7fffffffull) return (int)cudaErrorInvalidValue;
    const bool trace = job.trace_cap > 0u;
    const void *fn = trace ? (const void *)static_trial_kernel<ModelT, NPROC, NQUEUE, NEVENT, true>
                           : (const void *)static_trial_kernel<ModelT, NPROC, NQUEUE, NEVENT, false>;
    void *kargs[] = { (void *)&sa };
    cudaError_t e = cudaLaunchKernel(fn, dim3((unsigned)blocks), dim3(STATIC_BLOCK), kargs, 0, stream);
    if (e != cudaSuccess) return (int)e;
    e = cudaGetLastError();
    if (e != cudaSuccess || job.status == nullptr) return (int)e;      // nobody could see a flag: nothing to repair by
    return launch_model<ModelT<Sim>>(job, (unsigned char *)job.workspace + rings, job.workspace_bytes - rings, STATIC_REPAIR_BITS, stream);
}

}  // namespace cmb
}  // namespace cimba_b200

// A model template of the static tier in a library of its own: ModelT<cmb::StaticSim<NPROC, NQUEUE>> runs first, and
// ModelT<cmb::Sim> re-runs what that flags
#define CMB_EXPORT_STATIC_MODEL(ModelT, NPROC, NQUEUE, name_string) CMB_EXPORT_STATIC_MODEL_EVENTS(ModelT, NPROC, NQUEUE, 0, name_string)
// ... with NEVENT slots for events of the model's own (cmb_event_schedule), which also makes the event list order by priority
#define CMB_EXPORT_STATIC_MODEL_EVENTS(ModelT, NPROC, NQUEUE, NEVENT, name_string)                                    \
    extern "C" const char *cimba_b200_user_model_name(void) { return name_string; }                                  \
    extern "C" uint64_t cimba_b200_user_model_workspace_bytes(const cimba_b200_device_job *job)                      \
    {                                                                                                                 \
        return job ? cimba_b200::cmb::workspace_bytes_static<ModelT, NPROC, NQUEUE, NEVENT>(*job) : 0u;               \
    }                                                                                                                 \
    extern "C" int cimba_b200_user_model_launch(const cimba_b200_device_job *job, void *stream)                      \
    {                                                                                                                 \
        if (job == nullptr || job->workspace == nullptr) return (int)cudaErrorInvalidValue;                          \
        return cimba_b200::cmb::launch_static_model<ModelT, NPROC, NQUEUE, NEVENT>(*job, (cudaStream_t)stream);       \
    }

// A model in a library of its own: the three C entry points cimba_b200_model_load() looks up.
#define CMB_EXPORT_MODEL(Model, name_string)                                                                          \
    extern "C" const char *cimba_b200_user_model_name(void) { return name_string; }                                  \
    extern "C" uint64_t cimba_b200_user_model_workspace_bytes(const cimba_b200_device_job *job)                      \
    {                                                                                                                 \
        return job ? cimba_b200::cmb::workspace_bytes_for<Model>(*job) : 0u;                                          \
    }                                                                                                                 \
    extern "C" int cimba_b200_user_model_launch(const cimba_b200_device_job *job, void *stream)                      \
    {                                                                                                                 \
        if (job == nullptr || job->workspace == nullptr) return (int)cudaErrorInvalidValue;                          \
        return cimba_b200::cmb::launch_model<Model>(*job, (unsigned char *)job->workspace, job->workspace_bytes, 0u,  \
                                                    (cudaStream_t)stream);                                            \
    }
