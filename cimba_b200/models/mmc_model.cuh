// mmc_model.cuh - BASELINE config 3: M/M/c through cmb_resourcepool, a generator process
// that starts one customer PROCESS per arrival, finished customers recycled through a free list.  The oracle is the
// same model written against the reference's API, oracle/ref_build/ref_driver.c c_source_body / c_customer_body.
#pragma once
#include "../csrc/cmb_kernel.cuh"

namespace cimba_b200 {
namespace models {

struct MMC {
    cmb::resourcepool pool;
    uint32_t source, free_list, created;
    double   arr_mean, srv_mean;
    uint64_t num_objects, objects, ui;
    double   sum_wait;
    enum : uint32_t { SOURCE, CUSTOMER };

    // growth memory of one trial: a process record per customer alive at once (a few dozen at rho = 0.8, more in
    // overload), the wait list, an event list of servers + 2 entries - and as much again for what doubling leaves behind
    static uint64_t arena_bytes_per_trial(const cimba_b200_device_job &job)
    {
        return 65536u + (uint64_t)(job.servers > 0 ? job.servers : 1) * 512u;
    }

    CMB_FN void customer(cmb::Sim &sim, uint32_t me, int64_t sig)
    {
        MMC &m = *this;
        CMB_PROCESS_BEGIN
        CMB_RESOURCEPOOL_ACQUIRE(pool, 1u);
        CMB_PROCESS_HOLD_EXPONENTIAL(srv_mean);
        CMB_RESOURCEPOOL_RELEASE(pool, 1u);
        sum_wait += cmb_time() - sim.proc[me].f[0];                 // f[0] = the customer's arrival time
        objects += 1u;
        sim.proc[me].u[0] = free_list;                              // push self on the free list, then return
        free_list = me;
        CMB_PROCESS_END
    }

    CMB_FN void generator(cmb::Sim &sim, uint32_t me, int64_t sig)
    {
        MMC &m = *this;
        CMB_PROCESS_BEGIN
        for (ui = 0u; ui < num_objects; ui++) {
            CMB_PROCESS_HOLD_EXPONENTIAL(arr_mean);
            {
                uint32_t cu = free_list;
                if (cu != cmb::NIL) {
                    free_list = (uint32_t)sim.proc[cu].u[0];
                }
                else {
                    cu = cmb_process_create(CUSTOMER, 0, 0u);
                    created++;
                }
                sim.proc[cu].f[0] = cmb_time();
                cmb_process_start(cu);
            }
        }
        CMB_PROCESS_END
    }

    CMB_FN void run_trial(cmb::Sim &sim, const cmb::TrialIn &in)
    {
        arr_mean = in.arr_mean;
        srv_mean = in.srv_mean;
        num_objects = in.num_objects;
        objects = 0u;
        sum_wait = 0.0;
        free_list = cmb::NIL;
        created = 0u;
        cmb_resourcepool_initialize(pool, (uint64_t)in.servers);
        source = cmb_process_create(SOURCE, 0, 0u);
        cmb_process_start(source);
    }

    CMB_FN void process(cmb::Sim &sim, uint32_t me, uint32_t kind, int64_t sig)
    {
        if (kind == SOURCE) generator(sim, me, sig);
        else customer(sim, me, sig);
    }
    CMB_FN void event(cmb::Sim &, uint32_t, uint32_t, int64_t) {}
    CMB_FN bool demand(cmb::Sim &, uint32_t, uint32_t, int32_t) { return false; }

    CMB_FN void finish(cmb::Sim &, cmb::TrialOut &out)
    {
        out.objects = objects;
        out.sum_wait = sum_wait;
        out.max_queue = created;                        // process structs ever created (the oracle's all_count)
    }
};

}  // namespace models
}  // namespace cimba_b200
