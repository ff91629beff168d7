// pool_model.cuh - M/M/c through a counting-semaphore resource pool
// (cmb_resourcepool), one persistent simulation kernel, lane per trial.
//
// Reference model (no reference file exists for this workload; it is written once
// against the reference API in oracle/ref_build/ref_driver.c, model 2, following
// the process-per-customer pattern of tutorial/tut_4_1.c:319-339):
//   generator: loop N times { hold(exp(arr_mean)); c = recycled-or-new customer;
//                             c.t_arrival = cmb_time(); cmb_process_start(c) }
//   customer:  cmb_resourcepool_acquire(pool, 1); hold(exp(srv_mean));
//              cmb_resourcepool_release(pool, 1); sum += cmb_time() - t_arrival; exit
//
// Reference mechanics reproduced:
//   * cmb_process_start  -> START event at now (src/cmb_process.c:127-135); a
//     FINISHED process struct may be started again (src/cmi_coroutine.c:165-196).
//   * acquire (src/cmb_resourcepool.c:362-533, amount 1, no pre-emption):
//       loop { if (capacity - in_use >= 1) { in_use++; signal(guard); return }
//              wait(guard) }
//     signal wakes AT MOST the head waiter, and only if its demand (any unit
//     available, :198-211) holds, by scheduling WAKE_RESOURCE at now
//     (src/cmb_resourceguard.c:202-226); the woken process re-tests.
//   * release (:561-605): in_use--; signal(guard).
//   * the guard's wait list is ordered (priority desc, entry time asc, sequence
//     asc) (src/cmb_resourceguard.c:71-90).  Every process here has priority 0, so
//     that order is arrival order and the wait list is a FIFO.
//
// Device formulation: a customer is never in two places - it is either an entry
// of the future-event list (START / WAKE_TIME / WAKE_RESOURCE pending) or an
// entry of the guard's wait list - so its process-local state (the arrival time)
// travels with that entry instead of living in a process table: no per-trial
// process array, no slot allocator, and the whole trial state stays on chip:
//   EventList<16>  in shared memory (<= capacity + 2 live events),
//   wait list      = StampRing (32-entry shared-memory window, HBM spill ring),
//   RNG, clock, pool counters in registers.
#pragma once

#include "engine.cuh"
#include "rng.cuh"

namespace cimba_b200 {

constexpr int POOL_BLOCK = 64;
constexpr int POOL_FEL_CAP = 16;
constexpr int POOL_WINDOW = 32;
constexpr int POOL_COLD_BATCH = 4;

struct PoolArgs {
    int32_t  servers;
    uint64_t master_seed, first_trial, num_trials, num_objects;
    const double *arr_mean, *srv_mean;
    uint64_t *events, *objects;
    double   *t_end, *sum_wait;
    uint32_t *status, *max_queue;
    double   *spill;                   // [num_trials][spill_cap] wait-list spill
    uint32_t  spill_cap;
    uint64_t  trace_cap;
    uint64_t *trace_key;
    double   *trace_time;
    unsigned long long *diag;          // optional: [0] += event-loop iterations of each warp, [1] += warps (bench.py)
};

enum : uint32_t { TAG_GENERATOR = 0u, TAG_CUSTOMER = 1u };

template <bool TRACE>
__global__ void __launch_bounds__(POOL_BLOCK)
pool_kernel(const PoolArgs a)
{
    __shared__ ZigHot hot;
    __shared__ EventHead fel_head[POOL_FEL_CAP * POOL_BLOCK];
    __shared__ double fel_pay[POOL_FEL_CAP * POOL_BLOCK];
    __shared__ double wait_smem[POOL_WINDOW * POOL_BLOCK];

    stage_zig_hot(hot, false);
    __syncthreads();

    constexpr unsigned FULL = 0xffffffffu;
    const uint64_t trial = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool alive = trial < a.num_trials;

    Sfc64 rng;
    rng.a = rng.b = rng.c = rng.d = 0u;
    EventList<POOL_FEL_CAP> fel;
    StampRing<POOL_WINDOW> waiting;                    // arrival stamps of the customers queued at the guard
    fel.init(&fel_head[threadIdx.x], &fel_pay[threadIdx.x], POOL_BLOCK);
    waiting.init(&wait_smem[threadIdx.x], POOL_BLOCK,
                 (a.spill_cap && alive) ? a.spill + trial * a.spill_cap : nullptr, a.spill_cap);

    double now = 0.0, sum_wait = 0.0, arr_mean = 1.0, srv_mean = 1.0;
    uint32_t pops = 0u, produced = 0u, served = 0u, status = TRIAL_OK;
    uint32_t in_use = 0u, live = 0u, most_live = 0u, deepest = 0u;
    const uint32_t capacity = (uint32_t)a.servers;
    const uint32_t quota = (uint32_t)a.num_objects;

    if (alive) {
        arr_mean = a.arr_mean[trial];
        srv_mean = a.srv_mean[trial];
        rng.seed(fmix64(a.master_seed, a.first_trial + trial));
        fel.schedule(ACT_START, TAG_GENERATOR, 0.0, 0.0);       // cmb_process_start(source)
    }

    bool parked = false;               // waiting for company on the ziggurat slow path
    uint64_t parked_u = 0u;
    uint32_t parked_tag = 0u;
    double parked_mean = 0.0, parked_pay = 0.0;

    for (;;) {
        if (!__any_sync(FULL, alive)) {
            break;
        }
        // uniform trip count for the event-list scan
        const uint32_t scan = __reduce_max_sync(FULL, (alive && !parked) ? fel.count : 0u);

        bool draw = false;
        double draw_mean = 0.0, draw_pay = 0.0;
        uint32_t draw_tag = 0u;

        if (alive && !parked) {
            deepest = max(deepest, fel.count);
            EventHead ev;
            double pay;
            if (!fel.pop(scan, ev, pay)) {
                alive = false;                          // cmb_event_queue_execute returns
                if (a.events)    a.events[trial] = pops;
                if (a.objects)   a.objects[trial] = served;
                if (a.t_end)     a.t_end[trial] = now;
                if (a.sum_wait)  a.sum_wait[trial] = sum_wait;
                if (a.status)    a.status[trial] = status | (fel.issued > 0x3ffffff0u ? TRIAL_ERR_KEY_OVERFLOW : 0u);
                if (a.max_queue) a.max_queue[trial] = most_live;
            }
            else {
                now = ev.time;
                if (TRACE) {
                    if (pops < a.trace_cap) {
                        a.trace_key[trial * a.trace_cap + pops] = ev.keyact >> 2;
                        a.trace_time[trial * a.trace_cap + pops] = now;
                    }
                }
                pops++;
                const uint32_t act = ev.keyact & 3u;

                if (ev.tag == TAG_GENERATOR) {
                    if (act == ACT_WAKE_TIME) {
                        // back from hold: a customer arrives and is started
                        live++;
                        most_live = max(most_live, live);
                        if (!fel.schedule(ACT_START, TAG_CUSTOMER, now, now)) {
                            status |= TRIAL_ERR_FEL_OVERFLOW;
                        }
                        produced++;
                    }
                    if (produced < quota) {             // loop head: next inter-arrival hold
                        draw = true;
                        draw_mean = arr_mean;
                        draw_tag = TAG_GENERATOR;
                    }
                }
                else if (act == ACT_WAKE_TIME) {
                    // service finished: release, account, exit (the struct is recycled)
                    in_use--;
                    if (waiting.len > 0u && capacity - in_use > 0u) {       // guard signal
                        const double stamp = waiting.take();
                        if (!fel.schedule(ACT_WAKE_RESOURCE, TAG_CUSTOMER, now, stamp)) {
                            status |= TRIAL_ERR_FEL_OVERFLOW;
                        }
                    }
                    sum_wait = __dadd_rn(sum_wait, __dsub_rn(now, pay));
                    served++;
                    live--;
                }
                else {
                    // START or WAKE_RESOURCE: (re)try to acquire one unit
                    if (capacity - in_use >= 1u) {
                        in_use++;
                        // "in case someone else can use the leftovers" (:408-409)
                        if (waiting.len > 0u && capacity - in_use > 0u) {
                            const double stamp = waiting.take();
                            if (!fel.schedule(ACT_WAKE_RESOURCE, TAG_CUSTOMER, now, stamp)) {
                                status |= TRIAL_ERR_FEL_OVERFLOW;
                            }
                        }
                        draw = true;                    // hold(exp(srv_mean))
                        draw_mean = srv_mean;
                        draw_tag = TAG_CUSTOMER;
                        draw_pay = pay;
                    }
                    else if (!waiting.put(pay)) {       // cmb_resourceguard_wait, yield
                        status |= TRIAL_ERR_GUARD_OVERFLOW;
                    }
                }
            }
        }

        // ---- converged: draw the hold time, insert the wake-up
        if (draw) {
            const uint64_t u = rng.next();
            if (Sfc64::exp_is_hot(u)) {
                const double dur = __dmul_rn(draw_mean, Sfc64::exp_hot(hot, u));
                if (!fel.schedule(ACT_WAKE_TIME, draw_tag, __dadd_rn(now, dur), draw_pay)) {
                    status |= TRIAL_ERR_FEL_OVERFLOW;
                }
            }
            else {
                parked = true;
                parked_u = u;
                parked_tag = draw_tag;
                parked_mean = draw_mean;
                parked_pay = draw_pay;
            }
        }

        const unsigned pm = __ballot_sync(FULL, parked);
        if (pm != 0u) {
            const unsigned am = __ballot_sync(FULL, alive);
            if (__popc(pm) >= POOL_COLD_BATCH || pm == am) {
                if (parked) {
                    const double dur = __dmul_rn(parked_mean, rng.exp_cold(parked_u));
                    if (!fel.schedule(ACT_WAKE_TIME, parked_tag, __dadd_rn(now, dur), parked_pay)) {
                        status |= TRIAL_ERR_FEL_OVERFLOW;
                    }
                    parked = false;
                }
            }
        }
    }
    (void)deepest;
}

}  // namespace cimba_b200
