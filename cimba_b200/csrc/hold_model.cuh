// hold_model.cuh - the hold model: a deep future-event list, one trial per warp.
//
// Workload (oracle/ref_build/ref_driver.c model 7): `workers` processes that loop on
// cmb_process_hold(cmb_random_exponential(mean)), a ticker that holds exactly 1.0, and
// an end event at t = duration that stops everybody - the event-list shape of
// tutorial/tut_5_1.c (1000 target processes + a sensor ticking every second,
// SURVEY.md section 8d-5) without its float32 physics.  The list holds workers + 2
// entries for the whole run: every event is a dequeue and an enqueue on a heap
// ~1000 deep (src/cmi_hashheap.c:428-524, 10 binary levels of 64-byte tags there).
//
// This is the case BASELINE.json's north_star describes: the event list is far too
// big for registers, so ONE WARP owns a trial and its 32 lanes do the heap work
// together:
//   * the list is a 32-ary implicit heap in shared memory (node i's children are
//     32i+1 .. 32i+32; 1057 entries need only 3 levels), structure-of-arrays so the
//     32 children of a node are one coalesced, conflict-free row per field;
//   * one sift level = each lane loads one child, three warp-wide REDUX minimum
//     reductions (time high word, time low word, key) + a ballot pick the first
//     child under the reference's order (time asc, key asc; non-negative doubles
//     order like their bit patterns), lane 0 moves it up;
//   * dequeue-then-enqueue of an event step is fused into one replace-root +
//     sift-down (same resulting set, and pop order depends only on the set);
//   * the scalar part (clock, counters, the sfc64/ziggurat draw) is computed
//     redundantly by all 32 lanes, so nothing needs broadcasting.
// Any heap that realises the order pops in the reference's order (SURVEY.md
// section 9, "Event ordering"), so the arity is free.
#pragma once

#include "engine.cuh"
#include "rng.cuh"

namespace cimba_b200 {

constexpr int HOLD_CAP = 1088;                 // entries per trial: 1 + 32 + 1024 and slack
constexpr uint32_t HOLD_SMEM_BYTES = HOLD_CAP * 16u;

struct HoldArgs {
    int32_t  workers;
    uint64_t master_seed, first_trial, num_trials, duration;
    const double *mean;
    uint64_t *events, *objects;
    double   *t_end, *sum_wait;
    uint32_t *status, *max_queue;
    uint64_t *counters;
    uint64_t  trace_cap;
    uint64_t *trace_key;
    double   *trace_time;
};

// shared-window accessors on 32-bit addresses (no generic->shared conversion per access)
__device__ __forceinline__ unsigned long long lds_u64(uint32_t addr)
{
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u64(uint32_t addr, unsigned long long v)
{
    asm volatile("st.shared.u64 [%0], %1;" :: "r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
}

struct WarpHeap {
    uint32_t  t;               // shared-window byte address of time[] (IEEE-754 bit patterns; times are >= 0)
    uint32_t  key;             // ... of key[]
    uint32_t  info;            // ... of info[]: (process << 3) | action
    uint32_t  count;           // warp-uniform

    // Sink the entry (mt, mk, mi) from node i to its place (heap_down, src/cmi_hashheap.c:321-370,
    // generalised to 32 children per node).  All 32 lanes call this with identical arguments.
    __device__ __forceinline__ void sift_down(uint32_t i, unsigned long long mt, uint32_t mk, uint32_t mi)
    {
        constexpr unsigned FULL = 0xffffffffu;
        const unsigned lane = threadIdx.x & 31u;
        for (;;) {
            const uint32_t first = 32u * i + 1u;
            if (first >= count) {
                break;
            }
            const uint32_t c = first + lane;
            const bool valid = c < count;
            const unsigned long long ct = valid ? lds_u64(t + c * 8u) : ~0ull;
            const uint32_t ck = valid ? lds_u32(key + c * 4u) : 0xffffffffu;
            const uint32_t hi = (uint32_t)(ct >> 32), lo = (uint32_t)ct;
            // first child under (time asc, key asc).  Fast path: the high words of the
            // 32 times are almost always distinct, so one REDUX + one ballot decide.
            const uint32_t mhi = __reduce_min_sync(FULL, hi);
            unsigned cand = __ballot_sync(FULL, hi == mhi);
            uint32_t mlo = __shfl_sync(FULL, lo, __ffs(cand) - 1u);
            uint32_t mkey = __shfl_sync(FULL, ck, __ffs(cand) - 1u);
            if (__popc(cand) > 1) {                     // ties in the high word: settle low word, then key
                mlo = __reduce_min_sync(FULL, hi == mhi ? lo : 0xffffffffu);
                const bool tie = (hi == mhi) & (lo == mlo);
                mkey = __reduce_min_sync(FULL, tie ? ck : 0xffffffffu);
                cand = __ballot_sync(FULL, tie & (ck == mkey));
            }
            const unsigned long long best = ((unsigned long long)mhi << 32) | mlo;
            if (mt < best || (mt == best && mk < mkey)) {
                break;                                  // the moving entry goes before every child
            }
            const unsigned winner = __ffs(cand) - 1u;
            const uint32_t w = first + winner;
            if (lane == 0u) {                           // move the first child up
                sts_u64(t + i * 8u, best);
                sts_u32(key + i * 4u, mkey);
                sts_u32(info + i * 4u, lds_u32(info + w * 4u));
            }
            i = w;
            // the next round's store goes to a slot this round read; the warp-wide REDUX above already
            // orders them (it consumes the loaded values), this makes the ordering explicit
            __syncwarp();
        }
        if (lane == 0u) {
            sts_u64(t + i * 8u, mt);
            sts_u32(key + i * 4u, mk);
            sts_u32(info + i * 4u, mi);
        }
        __syncwarp();
    }

    // cmi_hashheap_enqueue: append and float up (heap_up, :277-316).  Scalar work: lane 0.
    __device__ __forceinline__ void push(unsigned long long mt, uint32_t mk, uint32_t mi)
    {
        if ((threadIdx.x & 31u) == 0u) {
            uint32_t i = count;
            while (i > 0u) {
                const uint32_t p = (i - 1u) >> 5;
                const unsigned long long pt = lds_u64(t + p * 8u);
                const uint32_t pk = lds_u32(key + p * 4u);
                if (!(mt < pt || (mt == pt && mk < pk))) {
                    break;
                }
                sts_u64(t + i * 8u, pt);
                sts_u32(key + i * 4u, pk);
                sts_u32(info + i * 4u, lds_u32(info + p * 4u));
                i = p;
            }
            sts_u64(t + i * 8u, mt);
            sts_u32(key + i * 4u, mk);
            sts_u32(info + i * 4u, mi);
        }
        count++;
        __syncwarp();
    }
};

template <bool TRACE>
__global__ void __launch_bounds__(32)
hold_kernel(const HoldArgs a)
{
    extern __shared__ __align__(16) unsigned char hold_smem[];
    WarpHeap h;
    h.t = (uint32_t)__cvta_generic_to_shared(hold_smem);
    h.key = h.t + HOLD_CAP * 8u;
    h.info = h.t + HOLD_CAP * 12u;
    h.count = 0u;
    asm volatile("" : "+r"(h.t), "+r"(h.key), "+r"(h.info));

    const unsigned lane = threadIdx.x & 31u;
    const uint32_t ticker = (uint32_t)a.workers;       // process index of the ticker

    // persistent CTA (one warp): trials blockIdx.x, blockIdx.x + gridDim.x, ...
    for (uint64_t trial = blockIdx.x; trial < a.num_trials; trial += gridDim.x) {
        Sfc64 rng;
        rng.seed(fmix64(a.master_seed, a.first_trial + trial));
        const double mean = a.mean[trial];
        double now = 0.0, sum_wait = 0.0;
        uint64_t pops = 0u, wakes = 0u, ticks = 0u;
        uint32_t issued = 0u, deepest = 0u;
        // one variate of look-ahead (the draw does not depend on the event list, so it
        // is taken off the pop -> push critical path; stream order is unchanged)
        double e_next = rng.std_exponential_global();

        // cmb_process_start for workers 0..n-1 and the ticker: START events at t = 0 with
        // keys 1, 2, ...  Equal times and ascending keys in index order already form a heap.
        __syncwarp();
        for (uint32_t j = lane; j <= ticker; j += 32u) {
            sts_u64(h.t + j * 8u, 0ull);
            sts_u32(h.key + j * 4u, j + 1u);
            sts_u32(h.info + j * 4u, (j << 3) | ACT_START);
        }
        h.count = ticker + 1u;
        issued = ticker + 1u;
        __syncwarp();
        h.push((unsigned long long)__double_as_longlong((double)a.duration), ++issued, (0xffffu << 3) | 5u);

        for (;;) {
            deepest = max(deepest, h.count);
            if (h.count == 0u) {
                break;
            }
            // cmi_hashheap_dequeue: the root (all lanes read it: a broadcast)
            const unsigned long long rt = lds_u64(h.t);
            const uint32_t rk = lds_u32(h.key);
            const uint32_t ri = lds_u32(h.info);
            __syncwarp();                               // every lane has the root before lane 0 overwrites it below
            now = __longlong_as_double((long long)rt);
            if (TRACE) {
                if (lane == 0u && pops < a.trace_cap) {
                    a.trace_key[trial * a.trace_cap + pops] = rk;
                    a.trace_time[trial * a.trace_cap + pops] = now;
                }
            }
            pops++;
            const uint32_t act = ri & 7u, who = ri >> 3;
            if (act == 5u) {
                // the end event stops every process; each owns exactly one pending hold, and
                // cmb_process_stop cancels it (src/cmb_process.c:698-723): the list is empty
                h.count = 0u;
                continue;
            }
            if (act == ACT_WAKE_TIME) {
                if (who == ticker) {
                    ticks++;
                }
                else {
                    wakes++;
                    sum_wait = __dadd_rn(sum_wait, now);
                }
            }
            // back in the body: hold again (cmb_process_hold -> cmb_event_schedule)
            const bool draws = who != ticker;
            const double when = __dadd_rn(now, draws ? __dmul_rn(mean, e_next) : 1.0);
            // dequeue + enqueue fused: the new event replaces the root and sinks
            h.sift_down(0u, (unsigned long long)__double_as_longlong(when), ++issued, (who << 3) | ACT_WAKE_TIME);
            if (draws) {
                e_next = rng.std_exponential_global();  // refill after use, overlapping the next pop
            }
        }

        if (lane == 0u) {
            if (a.events)    a.events[trial] = pops;
            if (a.objects)   a.objects[trial] = wakes;
            if (a.t_end)     a.t_end[trial] = now;
            if (a.sum_wait)  a.sum_wait[trial] = sum_wait;
            if (a.status)    a.status[trial] = TRIAL_OK;
            if (a.max_queue) a.max_queue[trial] = deepest;
            if (a.counters) {
                a.counters[trial * 8u + 0u] = wakes;
                a.counters[trial * 8u + 1u] = ticks;
                for (int k = 2; k < 8; k++) {
                    a.counters[trial * 8u + k] = 0u;
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace cimba_b200
