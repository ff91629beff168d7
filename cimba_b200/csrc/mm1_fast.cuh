// mm1_fast.cuh - the M/M/1 replication kernel, written for SIMT issue efficiency.
//
// Same model, same per-trial arithmetic and the same order of random draws, key
// issues and pops as queue_kernel<0> in queue_model.cuh (which stays as the
// readable formulation and serves G/G/1); see that file for the mapping onto
// benchmark/MM1_multi.c:52-125 and the reference's dispatcher.
//
// What is different is only how a warp executes it.  Profiling queue_kernel<0>
// on B200 (profiles/r01_mm1_baseline.md) showed 259 issued warp-instructions per
// event step at 21 of 32 lanes active: the arrival and service process bodies
// were separate divergent regions wrapped in reconvergence barriers, and 40 % of
// the issued instructions were register moves merging those regions' results.
// With ~3 warps per scheduler the kernel is bound by dependent-issue latency, so
// every issued instruction costs ~5 cycles.  Here one event step is a single
// straight-line, predicated sequence that all 32 lanes execute together:
//   pop-min over the two process-owned event slots (compare + selects),
//   the arrival body and the service body as predicated updates of the same
//   registers (a lane is in exactly one of them), the variate draw and the
//   wake-up insert.  The only branches left are the rare ones: queue spill to
//   HBM, trial completion and the parked ziggurat slow path.
#pragma once

#include "engine.cuh"
#include "queue_model.cuh"
#include "rng.cuh"

namespace cimba_b200 {

// 32-bit shared-window accesses: one address register, no generic->shared
// conversion per access (the static-__shared__ form costs five extra
// instructions each time on sm_100a).
__device__ __forceinline__ void sts_f64(uint32_t addr, double v)
{
    asm volatile("st.shared.f64 [%0], %1;" :: "r"(addr), "d"(v) : "memory");
}

__device__ __forceinline__ double lds_f64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}

template <bool TRACE>
__global__ void __launch_bounds__(QUEUE_BLOCK)
mm1_kernel(const QueueArgs a)
{
    __shared__ double exp_x[256];                       // ziggurat layer widths (hot table)
    __shared__ double ring_smem[QUEUE_WINDOW * QUEUE_BLOCK];

    for (unsigned i = threadIdx.x; i < 256u; i += blockDim.x) {
        exp_x[i] = zig::zig_exp_x[i];
    }
    __syncthreads();

    constexpr unsigned FULL = 0xffffffffu;
    constexpr uint32_t WMASK = QUEUE_WINDOW - 1;
    constexpr uint32_t ROW = QUEUE_BLOCK * 8u;          // bytes between consecutive ring entries of one trial
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
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

    // ---- per-trial state (registers)
    Sfc64 rng;
    rng.a = rng.b = rng.c = rng.d = 0u;
    // event slot 0 = arrival process, slot 1 = service process (SlotFel<2> spelled out)
    double t_arr = INF, t_srv = INF;
    uint32_t k_arr = 0u, k_srv = 0u, a_arr = ACT_NONE, a_srv = ACT_NONE;
    uint32_t issued = 0u;                               // item_counter, src/cmi_hashheap.c:449-453
    double now = 0.0, stamp = 0.0, sum_wait = 0.0;
    double arr_mean = 1.0, srv_mean = 1.0;
    uint32_t pops = 0u, produced = 0u, served = 0u, status = TRIAL_OK, longest = 0u;
    uint32_t q_head = 0u, q_len = 0u;
    bool server_waiting = false;
    const uint32_t quota = (uint32_t)a.num_objects;
    const uint32_t win = (uint32_t)__cvta_generic_to_shared(&ring_smem[threadIdx.x]);
    const uint32_t tab = (uint32_t)__cvta_generic_to_shared(&exp_x[0]);
    double *const spill = (a.spill_cap && alive) ? a.spill + trial * a.spill_cap : nullptr;
    const uint32_t spill_mask = a.spill_cap - 1u;

    if (alive) {
        arr_mean = a.arr_mean[trial];
        srv_mean = a.srv_mean[trial];
        rng.seed(fmix64(a.master_seed, a.first_trial + trial));
        t_arr = 0.0; k_arr = 1u; a_arr = ACT_START;     // cmb_process_start(arrival), MM1_multi.c:107-108
        t_srv = 0.0; k_srv = 2u; a_srv = ACT_START;     // cmb_process_start(service), :109-111
        issued = 2u;
    }

    bool parked = false;
    uint64_t parked_u = 0u;
    bool parked_is_arr = false;

    while (__any_sync(FULL, alive)) {
        // ---------------- speculative variate: the next sfc64 output and its
        // hot-path exponential, computed before we know whether this step draws
        // (independent of the pop, so it overlaps it); state is committed below.
        Sfc64 nxt = rng;
        const uint64_t u = nxt.next();
        const bool hot = Sfc64::exp_is_hot(u);
        const double std_exp = __dmul_rn(lds_f64(tab + ((uint32_t)u & 0xffu) * 8u), __ull2double_rn(u));

        // ---------------- pop-min (cmi_hashheap_dequeue order: time asc, key asc)
        const bool go0 = alive & !parked;
        const bool first_arr = (t_arr < t_srv) | ((t_arr == t_srv) & (k_arr < k_srv));
        const uint32_t act = first_arr ? a_arr : a_srv;
        const bool go = go0 & (act != ACT_NONE);
        const bool done = go0 & (act == ACT_NONE);      // event list ran dry
        const bool is_arr = go & first_arr;
        const bool is_srv = go & !first_arr;
        const bool wake = act == ACT_WAKE_TIME;
        if (TRACE) {
            if (go && pops < a.trace_cap) {
                a.trace_key[trial * a.trace_cap + pops] = first_arr ? k_arr : k_srv;
                a.trace_time[trial * a.trace_cap + pops] = first_arr ? t_arr : t_srv;
            }
        }
        now = go ? (first_arr ? t_arr : t_srv) : now;   // src/cmb_event.c:239-241
        pops += go ? 1u : 0u;
        t_arr = is_arr ? INF : t_arr;
        a_arr = is_arr ? ACT_NONE : a_arr;
        t_srv = is_srv ? INF : t_srv;
        a_srv = is_srv ? ACT_NONE : a_srv;

        // ---------------- arrival body (MM1_multi.c:58-66): back from hold -> put
        const bool put = is_arr & wake;
        const bool put_far = put & (q_len >= (uint32_t)QUEUE_WINDOW);
        if (put & !put_far) {
            sts_f64(win + ((q_head + q_len) & WMASK) * ROW, now);
        }
        if (put_far) {                                  // rare: beyond the on-chip window
            const uint32_t pos = q_head + q_len;
            if (spill != nullptr && q_len - QUEUE_WINDOW <= spill_mask) {
                spill[pos & spill_mask] = now;
            }
            else {
                status |= TRIAL_ERR_QUEUE_OVERFLOW;
                q_len--;                                // entry dropped
            }
        }
        q_len += put ? 1u : 0u;
        produced += put ? 1u : 0u;
        longest = max(longest, q_len);
        // cmb_objectqueue_put -> cmb_resourceguard_signal(front guard): wake the server
        const bool ring_bell = put & server_waiting;
        issued += ring_bell ? 1u : 0u;
        t_srv = ring_bell ? now : t_srv;
        k_srv = ring_bell ? issued : k_srv;
        a_srv = ring_bell ? ACT_WAKE_RESOURCE : a_srv;
        server_waiting = server_waiting & !ring_bell;

        // ---------------- service body (MM1_multi.c:78-88)
        const bool finished = is_srv & wake;            // back from the service hold
        const double new_sum = __dadd_rn(sum_wait, __dsub_rn(now, stamp));
        sum_wait = finished ? new_sum : sum_wait;
        served += finished ? 1u : 0u;
        // cmb_objectqueue_get: take the head or wait at the front guard
        const bool take = is_srv & (q_len > 0u);
        if (take) {
            const uint32_t slot = win + (q_head & WMASK) * ROW;
            stamp = lds_f64(slot);
            if (q_len > (uint32_t)QUEUE_WINDOW) {       // rare: refill the freed slot from HBM
                sts_f64(slot, spill[(q_head + QUEUE_WINDOW) & spill_mask]);
            }
        }
        q_head += take ? 1u : 0u;
        q_len -= take ? 1u : 0u;
        server_waiting = server_waiting | (is_srv & !take);

        // ---------------- hold: commit the draw, insert the wake-up
        const bool draw = take | (is_arr & (produced < quota));
        const bool push = draw & hot;
        const double when = __dadd_rn(now, __dmul_rn(is_arr ? arr_mean : srv_mean, std_exp));
        rng.a = draw ? nxt.a : rng.a;
        rng.b = draw ? nxt.b : rng.b;
        rng.c = draw ? nxt.c : rng.c;
        rng.d = draw ? nxt.d : rng.d;
        issued += push ? 1u : 0u;
        const bool push_arr = push & is_arr;
        const bool push_srv = push & is_srv;
        t_arr = push_arr ? when : t_arr;
        k_arr = push_arr ? issued : k_arr;
        a_arr = push_arr ? ACT_WAKE_TIME : a_arr;
        t_srv = push_srv ? when : t_srv;
        k_srv = push_srv ? issued : k_srv;
        a_srv = push_srv ? ACT_WAKE_TIME : a_srv;
        const bool park = draw & !hot;
        parked = parked | park;
        parked_u = park ? u : parked_u;
        parked_is_arr = park ? is_arr : parked_is_arr;

        // ---------------- rare paths
        if (done) {
            alive = false;
            if (a.events)    a.events[trial] = pops;
            if (a.objects)   a.objects[trial] = served;
            if (a.t_end)     a.t_end[trial] = now;
            if (a.sum_wait)  a.sum_wait[trial] = sum_wait;
            if (a.status)    a.status[trial] = status | (issued > 0xfffffff0u ? TRIAL_ERR_KEY_OVERFLOW : 0u);
            if (a.max_queue) a.max_queue[trial] = longest;
        }
        const unsigned pm = __ballot_sync(FULL, parked);
        if (pm != 0u) {
            const unsigned am = __ballot_sync(FULL, alive);
            if (__popc(pm) >= COLD_BATCH || pm == am) {
                if (parked) {
                    const double mean = parked_is_arr ? arr_mean : srv_mean;
                    const double dur = __dmul_rn(mean, rng.exp_cold(parked_u));
                    const double at = __dadd_rn(now, dur);
                    issued++;
                    if (parked_is_arr) { t_arr = at; k_arr = issued; a_arr = ACT_WAKE_TIME; }
                    else               { t_srv = at; k_srv = issued; a_srv = ACT_WAKE_TIME; }
                    parked = false;
                }
            }
        }
    }
}

}  // namespace cimba_b200
