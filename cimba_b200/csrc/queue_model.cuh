// queue_model.cuh - single-server queueing workloads (M/M/1, G/G/1) as one
// persistent simulation kernel.
//
// Reference model: benchmark/MM1_multi.c:52-125 - an arrival process
//     loop { hold(draw); stamp = cmb_time(); cmb_objectqueue_put(q, stamp) }
// and a service process
//     loop { cmb_objectqueue_get(q, &stamp); hold(draw); sum += cmb_time() - stamp }
// run by cmb_event_queue_execute() until the event list is empty.
//
// Device formulation: the two processes are resume-point machines that hand a
// "hold for a variate with mean m" command to the dispatcher; the dispatcher
// draws the variate and inserts the wake-up event in code that all lanes of the
// warp execute together, whichever process each lane's trial happens to be in.
// The ziggurat's 1.6 % slow path would otherwise be entered by some lane in 40 %
// of warp iterations; lanes that need it park (their trial simply does not
// advance) until a few of them can run it together.  Parking changes nothing a
// trial can observe: trials share no state, and within a trial the order of
// random draws, key issues and pops is untouched.
#pragma once

#include "engine.cuh"
#include "rng.cuh"
#include "summary.cuh"

namespace cimba_b200 {

constexpr int QUEUE_BLOCK = 64;        // threads per CTA: 7 CTAs/SM x 148 SMs >= 65 536 lanes in one wave
constexpr int QUEUE_WINDOW = 32;       // shared-memory ring entries per trial
constexpr int COLD_BATCH = 4;          // parked lanes needed before the slow path is run

struct QueueArgs {
    int32_t  mapping;                  // 1 = lane per trial, 32 = warp per trial
    uint64_t master_seed, first_trial, num_trials, num_objects;
    const double *arr_mean, *srv_mean;
    uint64_t *events, *objects;
    double   *t_end, *sum_wait;
    uint32_t *status, *max_queue;
    uint64_t *counters;                // [num_trials][8]: RECORD kernels write the queue-length cmb_wtdsummary here
    double   *spill;                   // [num_trials][spill_cap]
    uint32_t  spill_cap;               // power of two, or 0
    uint64_t  trace_cap;
    uint64_t *trace_key;
    double   *trace_time;
    unsigned long long *diag;          // optional: [0] += event-loop iterations of each warp, [1] += warps (bench.py)
};

enum { PROC_ARRIVAL = 0, PROC_SERVICE = 1 };

// RECORD = the queue's history is on (cmb_objectqueue_recording_start, src/cmb_objectqueue.c:161-177,
// as tutorial/tut_1_*.c and test/test_cimba.c run it): every put and get is a (length, time)
// sample of a cmb_timeseries, folded on the fly into the time-weighted cmb_wtdsummary that
// cmb_timeseries_summarize (src/cmb_timeseries.c:167-188) would compute from the stored history.
template <int MODEL, bool TRACE, bool RECORD = false>
__global__ void __launch_bounds__(QUEUE_BLOCK)
queue_kernel(const QueueArgs a)
{
    __shared__ ZigHot hot;
    __shared__ double ring_smem[QUEUE_WINDOW * QUEUE_BLOCK];

    stage_zig_hot(hot, MODEL == 1);
    __syncthreads();

    constexpr unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t trial;
    bool alive;
    if (a.mapping == 32) {
        trial = gtid >> 5;
        alive = (lane == 0u) && (trial < a.num_trials);
    }
    else {
        trial = gtid;
        alive = trial < a.num_trials;
    }

    // ---- per-trial state, all in registers
    Sfc64 rng;
    SlotFel<2> fel;
    StampRing<QUEUE_WINDOW> q;
    double now = 0.0, stamp = 0.0, sum_wait = 0.0;
    double arr_mean = 1.0, srv_mean = 1.0;
    uint64_t events = 0u;
    uint32_t produced = 0u, served = 0u, status = TRIAL_OK, longest = 0u;
    bool server_waiting = false;       // the front guard's only possible waiter
    const uint32_t quota = (uint32_t)a.num_objects;

    TimeWeighted hist;
    hist.start();
    if (RECORD) {
        hist.sample(0.0, 0.0);         // recording_start: the empty queue at t = 0
    }

    fel.clear();
    q.init(&ring_smem[threadIdx.x], QUEUE_BLOCK,
           (a.spill_cap && alive) ? a.spill + trial * a.spill_cap : nullptr, a.spill_cap);
    if (alive) {
        arr_mean = a.arr_mean[trial];
        srv_mean = a.srv_mean[trial];
        rng.seed(fmix64(a.master_seed, a.first_trial + trial));    // test/test_cimba.c:396
        fel.schedule(PROC_ARRIVAL, ACT_START, 0.0);                // benchmark/MM1_multi.c:107-108
        fel.schedule(PROC_SERVICE, ACT_START, 0.0);                // :109-111
    }

    bool parked = false;               // waiting for company on the ziggurat slow path
    uint64_t parked_u = 0u;
    int parked_who = 0;

    for (;;) {
        if (!__any_sync(FULL, alive)) {
            break;
        }

        bool draw = false;
        int who = 0;
        if (alive && !parked) {
            uint32_t act, key;
            double when;
            if (!fel.pop(who, act, when, key)) {
                // event list ran dry: cmb_event_queue_execute returns (src/cmb_event.c:259-267)
                alive = false;
                if (events > 0xfffffff0ull) {
                    status |= TRIAL_ERR_KEY_OVERFLOW;
                }
                if (a.events)    a.events[trial] = events;
                if (a.objects)   a.objects[trial] = served;
                if (a.t_end)     a.t_end[trial] = now;
                if (a.sum_wait)  a.sum_wait[trial] = sum_wait;
                if (a.status)    a.status[trial] = status;
                if (a.max_queue) a.max_queue[trial] = longest;
                if (RECORD) {
                    hist.sample((double)q.len, now);            // recording_stop closes the last interval
                    if (a.counters) wtd_store_row(hist.acc, a.counters + trial * 8u);
                }
            }
            else {
                now = when;            // src/cmb_event.c:239-241
                if (TRACE) {
                    if (events < a.trace_cap) {
                        a.trace_key[trial * a.trace_cap + events] = key;
                        a.trace_time[trial * a.trace_cap + events] = now;
                    }
                }
                events++;

                if (who == PROC_ARRIVAL) {
                    if (act == ACT_WAKE_TIME) {
                        // back from hold: stamp an object and put it (MM1_multi.c:61-64)
                        if (!q.put(now)) {
                            status |= TRIAL_ERR_QUEUE_OVERFLOW;
                        }
                        longest = max(longest, q.len);
                        if (RECORD) {
                            hist.sample((double)q.len, now);    // record_sample in put, src/cmb_objectqueue.c:289
                        }
                        produced++;
                        // cmb_objectqueue_put -> cmb_resourceguard_signal(front_guard):
                        // wake the head waiter if has_content holds (src/cmb_objectqueue.c:286-287,
                        // src/cmb_resourceguard.c:202-226)
                        if (server_waiting) {
                            server_waiting = false;
                            if (!fel.schedule(PROC_SERVICE, ACT_WAKE_RESOURCE, now)) {
                                status |= TRIAL_ERR_FEL_OVERFLOW;
                            }
                        }
                    }
                    // loop head (MM1_multi.c:58): next inter-arrival, or return -> cmb_process_exit
                    draw = produced < quota;
                }
                else {
                    if (act == ACT_WAKE_TIME) {
                        // service finished (MM1_multi.c:83-87)
                        sum_wait = __dadd_rn(sum_wait, __dsub_rn(now, stamp));
                        served++;
                    }
                    // cmb_objectqueue_get (src/cmb_objectqueue.c:203-260); after a resource
                    // wake-up the loop re-tests (:213)
                    if (q.len > 0u) {
                        stamp = q.take();
                        if (RECORD) {
                            hist.sample((double)q.len, now);    // ... and in get, :226-229
                        }
                        draw = true;
                    }
                    else {
                        server_waiting = true;      // cmb_resourceguard_wait, yield (:125-152)
                    }
                }
            }
        }

        // ---- converged: draw the hold time, insert the wake-up (cmb_process_hold,
        // src/cmb_process.c:262-285 -> cmb_event_schedule)
        if (draw) {
            const double mean = (who == PROC_ARRIVAL) ? arr_mean : srv_mean;
            if (MODEL == 0) {
                const uint64_t u = rng.next();
                if (Sfc64::exp_is_hot(u)) {
                    const double dur = __dmul_rn(mean, Sfc64::exp_hot(hot, u));
                    fel.schedule(who, ACT_WAKE_TIME, __dadd_rn(now, dur));
                }
                else {
                    parked = true;
                    parked_u = u;
                    parked_who = who;
                }
            }
            else {
                double dur;
                if (who == PROC_ARRIVAL) {
                    dur = rng.erlang(hot, 2u, __dmul_rn(0.5, mean));
                }
                else {
                    do {
                        dur = rng.normal(hot, mean, __dmul_rn(0.25, mean));
                    } while (dur < 0.0);
                }
                fel.schedule(who, ACT_WAKE_TIME, __dadd_rn(now, dur));
            }
        }

        if (MODEL == 0) {
            const unsigned pm = __ballot_sync(FULL, parked);
            if (pm != 0u) {
                const unsigned am = __ballot_sync(FULL, alive);
                if (__popc(pm) >= COLD_BATCH || pm == am) {
                    if (parked) {
                        const double mean = (parked_who == PROC_ARRIVAL) ? arr_mean : srv_mean;
                        const double dur = __dmul_rn(mean, rng.exp_cold(parked_u));
                        fel.schedule(parked_who, ACT_WAKE_TIME, __dadd_rn(now, dur));
                        parked = false;
                    }
                }
            }
        }
    }
}

}  // namespace cimba_b200
