// mm1_pc.cuh - M/M/1 with the variates produced by separate warps (job->variant = 2): an experiment on the headline kernel's
// binding resource, measured in profiles/r02_mm1.md.
//
// mm1_kernel (mm1_fast.cuh) is bound by dependent-issue latency at 3.5 warps per scheduler - all that 65 536 lane-trials give -
// with 62 % of the issue slots busy.  A standard exponential variate does not depend on the simulation (only which mean it is
// multiplied by does), so the random stream of a trial can be turned into variates by ANOTHER warp: here a CTA has two consumer
// warps running the event step of mm1_kernel without its sfc64 step, ziggurat look-up and slow-path parking, and two producer
// warps - lane for lane the same 64 trials - that keep a 16-entry ring of finished variates per trial in shared memory filled.
// Twice the warps per scheduler, the same work split in two: the question is whether the issue slots fill.
//
// Stream order is untouched: trial i's producer lane draws exactly the sequence mm1_kernel's lane would (first output, slow path
// with its extra draws where the ziggurat asks for one), the consumer takes variate k at its k-th hold.
// Hand-off: ring[k & 15] then prod_count, read as prod_count then ring; the consumer publishes how many it has taken every 2nd
// step, which is when the producer learns of room.  A lane whose ring looks empty skips the step.
//
// RESULT (profiles/r02_mm1.md): same answers, pop for pop; 1.29e11 events/s against mm1_kernel's 1.33e11.  The issue slots do
// fill further (62 % -> 68 %, 5.2 warps per scheduler), but the work is not split for free: 6.8 warp-instructions per event
// against 4.8 (the hand-off, the producers' polling, half-empty producer steps), and the half-rate ALU pipe - 70 % busy, 1.4 stall
// cycles of math-pipe throttle per issued instruction - is where both kinds of warp queue.  Kept as variant 2 for the record.
#pragma once

#include "mm1_fast.cuh"

namespace cimba_b200 {

constexpr int MM1_PC_CONSUMERS = 64;                    // trials per CTA; the CTA has as many producer threads again
constexpr uint32_t MM1_PC_DEPTH = 16u;
#ifndef MM1_PC_REFRESH_MASK
#define MM1_PC_REFRESH_MASK 1u     // the consumer publishes its count and looks at the producer's every 2nd step (measured: 8th 1.15e11, 4th 1.21e11, 2nd 1.28e11, every 1.29e11)
#endif
#ifndef MM1_PC_BURST
#define MM1_PC_BURST 8
#endif
#ifndef MM1_PC_SLEEP_NS
#define MM1_PC_SLEEP_NS 1000
#endif

// The counters of the hand-off.  -DMM1_PC_FENCED: ld.acquire / st.release at CTA scope (each a MEMBAR.CTA on sm_100a); default:
// volatile accesses, relying on a thread's shared-memory accesses being performed in program order by the SM's load-store path
// (which is what makes the ring entry visible before the count that announces it).  Measured both ways, profiles/r02_mm1.md.
__device__ __forceinline__ uint32_t ld_acquire_u32(uint32_t addr)
{
    uint32_t v;
#ifdef MM1_PC_FENCED
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
#else
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
#endif
    return v;
}

__device__ __forceinline__ void st_release_u32(uint32_t addr, uint32_t v)
{
#ifdef MM1_PC_FENCED
    asm volatile("st.release.cta.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
#else
    asm volatile("st.volatile.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
#endif
}

template <bool TRACE>
__global__ void __launch_bounds__(2 * MM1_PC_CONSUMERS, 7)
mm1_pc_kernel(const QueueArgs a)
{
    __shared__ double exp_x[256];
    __shared__ double ring_smem[QUEUE_WINDOW * MM1_PC_CONSUMERS];           // the object queue's windows (as mm1_kernel)
    __shared__ double var_smem[MM1_PC_DEPTH * MM1_PC_CONSUMERS];            // the variate rings, one column per trial
    __shared__ double scratch_smem[MM1_PC_CONSUMERS];
    __shared__ uint32_t prod_count[MM1_PC_CONSUMERS], cons_count[MM1_PC_CONSUMERS];
    __shared__ uint32_t done_flag[MM1_PC_CONSUMERS / 32];

    for (unsigned i = threadIdx.x; i < 256u; i += blockDim.x) exp_x[i] = zig::zig_exp_x[i];
    if (threadIdx.x < (unsigned)MM1_PC_CONSUMERS) {
        prod_count[threadIdx.x] = 0u;
        cons_count[threadIdx.x] = 0u;
    }
    if (threadIdx.x < (unsigned)(MM1_PC_CONSUMERS / 32)) done_flag[threadIdx.x] = 0u;
    __syncthreads();

    constexpr unsigned FULL = 0xffffffffu;
    const bool producer = threadIdx.x >= (unsigned)MM1_PC_CONSUMERS;
    const unsigned col = producer ? threadIdx.x - MM1_PC_CONSUMERS : threadIdx.x;      // which of the CTA's trials
    const uint64_t trial = (uint64_t)blockIdx.x * MM1_PC_CONSUMERS + col;
    const bool exists = trial < a.num_trials;
    const uint32_t var = (uint32_t)__cvta_generic_to_shared(&var_smem[col]);
    constexpr uint32_t VROW = MM1_PC_CONSUMERS * 8u;
    const uint32_t pc_addr = (uint32_t)__cvta_generic_to_shared(&prod_count[col]);
    const uint32_t cc_addr = (uint32_t)__cvta_generic_to_shared(&cons_count[col]);
    const uint32_t flag_addr = (uint32_t)__cvta_generic_to_shared(&done_flag[col >> 5]);
    const uint32_t tab = (uint32_t)__cvta_generic_to_shared(&exp_x[0]);

    if (producer) {
        // ---------------------------------------------------------------- the variate producer of trial `trial`
        Sfc64 rng;
        rng.a = rng.b = rng.c = rng.d = 0u;
        if (exists) rng.seed(fmix64(a.master_seed, a.first_trial + trial));
        uint32_t made = 0u, taken = 0u;
        bool parked = false;
        uint64_t parked_u = 0u;
        for (;;) {
            if (ld_acquire_u32(flag_addr) != 0u) break;                     // the consumer warp has finished all its trials
            taken = ld_acquire_u32(cc_addr);
            // a burst: the consumer says what it has taken every 8th step, so room comes in lumps of about eight
            bool any_made = false;
#pragma unroll 1
            for (int burst = 0; burst < MM1_PC_BURST; burst++) {
                taken = ld_acquire_u32(cc_addr);
                const bool room = exists & !parked & (made - taken < MM1_PC_DEPTH);
                if (!__any_sync(FULL, room)) break;
                any_made = true;
                if (room) {
                    const uint64_t u = rng.next();
                    if (Sfc64::exp_is_hot(u)) {
                        sts_f64(var + (made & (MM1_PC_DEPTH - 1u)) * VROW,
                                __dmul_rn(lds_f64(tab + ((uint32_t)u & 0xffu) * 8u), __ull2double_rn(u)));
                        made++;
                    }
                    else {
                        parked = true;
                        parked_u = u;
                    }
                }
            }
            const unsigned pm = __ballot_sync(FULL, parked);
            if (pm != 0u) {
                const unsigned starving = __ballot_sync(FULL, parked & (made - taken < 6u));
                if (__popc(pm) >= MM1_COLD_BATCH || starving != 0u) {
                    if (parked) {
                        sts_f64(var + (made & (MM1_PC_DEPTH - 1u)) * VROW, rng.exp_cold(parked_u));
                        made++;
                        parked = false;
                    }
                }
            }
            st_release_u32(pc_addr, made);
            if (!any_made) __nanosleep(MM1_PC_SLEEP_NS);                    // every ring full: leave the issue slots to the consumers
        }
        return;
    }

    // -------------------------------------------------------------------- the consumer: mm1_kernel's event step
    constexpr uint32_t WMASK = QUEUE_WINDOW - 1;
    constexpr uint32_t ROW = MM1_PC_CONSUMERS * 8u;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    double t_arr = INF, t_srv = INF;
    uint32_t k_arr = 0u, k_srv = 0u;
    uint32_t issued = 0u;
    double now = 0.0, stamp = 0.0, sum_wait = 0.0;
    double arr_mean = 1.0, srv_mean = 1.0;
    uint32_t produced = 0u, served = 0u, dropped = 0u, status = TRIAL_OK, longest = 0u;
    const uint32_t quota = (uint32_t)a.num_objects;
    uint32_t win = (uint32_t)__cvta_generic_to_shared(&ring_smem[col]);
    uint32_t scratch = (uint32_t)__cvta_generic_to_shared(&scratch_smem[col]);
    asm volatile("" : "+r"(win), "+r"(scratch));
    double *const spill = (a.spill_cap && exists) ? a.spill + trial * a.spill_cap : nullptr;
    const uint32_t spill_mask = a.spill_cap - 1u;
    uint32_t pops = 0u;
    uint32_t used = 0u, seen = 0u;                      // variates taken; the producer's count as last looked at

    if (exists) {
        arr_mean = a.arr_mean[trial];
        srv_mean = a.srv_mean[trial];
        t_arr = 0.0; k_arr = pack_key(1u, ACT_START);
        t_srv = 0.0; k_srv = pack_key(2u, ACT_START);
        issued = 2u;
    }
    bool alive = exists;
    uint32_t step = 0u;

    while (__any_sync(FULL, alive)) {
        // the next variate, fetched before it is known to be needed (and harmlessly if it is not there yet)
        const double e_next = lds_f64(var + (used & (MM1_PC_DEPTH - 1u)) * VROW);
        const bool go0 = alive & (used != seen);
        const bool first_arr = (t_arr < t_srv) | ((t_arr == t_srv) & (k_arr < k_srv));
        const uint32_t key = first_arr ? k_arr : k_srv;
        const uint32_t act = key & 3u;
        const bool go = go0 & (key != 0u);
        const bool done = go0 & (key == 0u);
        const bool is_arr = go & first_arr;
        const bool is_srv = go & !first_arr;
        const bool wake = act == ACT_WAKE_TIME;
        if (TRACE) {
            if (go && pops < a.trace_cap) {
                a.trace_key[trial * a.trace_cap + pops] = key >> 2;
                a.trace_time[trial * a.trace_cap + pops] = first_arr ? t_arr : t_srv;
            }
            pops += go ? 1u : 0u;
        }
        if (go) now = first_arr ? t_arr : t_srv;

        // arrival body
        const uint32_t q_len = produced - served;
        const bool put = is_arr & wake;
        const bool put_far = put & (q_len >= (uint32_t)QUEUE_WINDOW);
        sts_f64((put & !put_far) ? win + (produced & WMASK) * ROW : scratch, now);
        if (put_far) {
            if (spill != nullptr && q_len - QUEUE_WINDOW <= spill_mask) {
                spill[produced & spill_mask] = now;
            }
            else {
                status |= TRIAL_ERR_QUEUE_OVERFLOW;
                dropped++;
                served++;
            }
        }
        if (put) produced++;
        longest = max(longest, produced - served);
        const bool ring_bell = put & (k_srv == 0u);
        if (ring_bell) {
            issued++;
            t_srv = now;
            k_srv = pack_key(issued, ACT_WAKE_RESOURCE);
        }

        // service body
        const bool finished = is_srv & wake;
        const double new_sum = __dadd_rn(sum_wait, __dsub_rn(now, stamp));
        if (finished) sum_wait = new_sum;
        const bool take = is_srv & (produced != served);
        const uint32_t head_slot = win + (served & WMASK) * ROW;
        const double head_stamp = lds_f64(head_slot);
        if (take) stamp = head_stamp;
        if (take & (produced - served > (uint32_t)QUEUE_WINDOW)) {
            sts_f64(head_slot, spill[(served + QUEUE_WINDOW) & spill_mask]);
        }
        if (take) served++;

        // hold: take the variate, insert the wake-up
        const bool draw = take | (is_arr & (produced < quota));
        const double when = __dadd_rn(now, __dmul_rn(is_arr ? arr_mean : srv_mean, e_next));
        if (draw) {
            issued++;
            used++;
        }
        const double t_new = draw ? when : INF;
        const uint32_t k_new = draw ? pack_key(issued, ACT_WAKE_TIME) : 0u;
        if (is_arr) { t_arr = t_new; k_arr = k_new; }
        if (is_srv) { t_srv = t_new; k_srv = k_new; }

        if (done) alive = false;
        if ((++step & MM1_PC_REFRESH_MASK) == 0u || used == seen) {
            // look at what the producer has made, say what has been taken (the producer learns of room here)
            st_release_u32(cc_addr, used);
            seen = ld_acquire_u32(pc_addr);
        }
    }
    // every update above is predicated on `go`: the registers hold each trial's results
    if (exists) {
        if (a.events)    a.events[trial] = issued;
        if (a.objects)   a.objects[trial] = served - dropped;
        if (a.t_end)     a.t_end[trial] = now;
        if (a.sum_wait)  a.sum_wait[trial] = sum_wait;
        if (a.status)    a.status[trial] = status | (issued > 0x3ffffff0u ? TRIAL_ERR_KEY_OVERFLOW : 0u);
        if (a.max_queue) a.max_queue[trial] = longest;
    }
    __syncwarp();
    if ((threadIdx.x & 31u) == 0u) st_release_u32(flag_addr, 1u);          // this warp's producers may go
    if (a.diag != nullptr && (threadIdx.x & 31u) == 0u) {
        atomicAdd(a.diag, (unsigned long long)step);
        atomicAdd(a.diag + 1, 1ull);
    }
}

}  // namespace cimba_b200
