// hold_deep.cuh - the hold model with the future-event list's deep levels in HBM/L2.
//
// Same workload and same results as hold_model.cuh (ref_driver.c model 7: `workers`
// processes in cmb_process_hold(exp) loops, a ticker, an end event), but the event list
// is laid out the way BASELINE.json's north_star asks for a list too big for the chip:
// a warp-private 32-ary heap whose top lives on chip and whose deep levels are spilled
// to HBM and moved with coalesced 128-bit loads and stores:
//   level 0  the root                        replicated in every lane's registers
//   level 1  nodes 1..32                     one node per LANE, in registers
//   level 2  nodes 33..1056                  global memory, 16-byte records, row of 32
//   level 3  nodes 1057..33824               children = one 512-byte coalesced read
// One event = replace the root by the process's next wake-up and sink it:
//   * level 1: a warp-wide REDUX minimum over the lanes' registers, no memory at all;
//   * each deeper level: every lane loads one child of the chosen node (32 x 16 B,
//     one coalesced request, served by L2 when the working set fits its 126 MB and by
//     HBM otherwise), REDUX + ballot pick the first child, ONE lane stores 16 bytes.
// Algorithmic traffic: 512 B read + <= 32 B written per event and level below 1,
// against the 64-byte tags x 10 binary levels of src/cmi_hashheap.c:321-370.
//
// With no shared memory per trial the occupancy limit of hold_model.cuh (17 KB of
// shared memory per warp -> 12 warps per SM) is gone: 40+ warps per SM hide the L2
// latency of the row reads.  hold_model.cuh stays as job->variant = 1 (all in shared
// memory, <= 1080 workers) for comparison; this kernel takes up to 33 822 workers.
//
// Any heap that realises the order (time asc, key asc) pops in the reference's order
// (SURVEY.md section 9), so arity and placement are free; results are bit-identical.
#pragma once

#include "engine.cuh"
#include "hold_model.cuh"
#include "mm1_fast.cuh"
#include "rng.cuh"

namespace cimba_b200 {

constexpr int DEEP_BLOCK = 128;                         // 4 independent warps (trials) per CTA
constexpr uint32_t DEEP_MAX_ENTRIES = 1u + 32u + 1024u + 32768u;

struct DeepArgs {
    HoldArgs h;
    uint4   *rows;              // [resident warps][row_entries] level >= 2 records {time lo, time hi, key, info}
    uint64_t row_entries;       // per warp, multiple of 32
};

struct Picked {
    unsigned lane;
    unsigned long long t;
    uint32_t key;
};

// first of 32 candidates under (time asc, key asc); lanes without a candidate pass ~0
__device__ __forceinline__ Picked warp_first(unsigned long long ct, uint32_t ck)
{
    constexpr unsigned FULL = 0xffffffffu;
    const uint32_t hi = (uint32_t)(ct >> 32), lo = (uint32_t)ct;
    const uint32_t mhi = __reduce_min_sync(FULL, hi);
    unsigned cand = __ballot_sync(FULL, hi == mhi);
    if (__popc(cand) > 1) {                             // ties in the high word: settle low word, then key
        const uint32_t mlo = __reduce_min_sync(FULL, hi == mhi ? lo : 0xffffffffu);
        const bool tie = (hi == mhi) & (lo == mlo);
        const uint32_t mkey = __reduce_min_sync(FULL, tie ? ck : 0xffffffffu);
        cand = __ballot_sync(FULL, tie & (ck == mkey));
    }
    Picked p;
    p.lane = __ffs(cand) - 1u;
    p.t = __shfl_sync(FULL, ct, p.lane);
    p.key = __shfl_sync(FULL, ck, p.lane);
    return p;
}

__device__ __forceinline__ bool goes_before(unsigned long long at, uint32_t ak, unsigned long long bt, uint32_t bk)
{
    return at < bt || (at == bt && ak < bk);
}

// cmb_random_std_exponential with the layer table at a 32-bit shared-window address
__device__ __forceinline__ double std_exponential_shared(Sfc64 &rng, uint32_t tab)
{
    const uint64_t u = rng.next();
    return Sfc64::exp_is_hot(u)
        ? __dmul_rn(lds_f64(tab + ((uint32_t)u & 0xffu) * 8u), __ull2double_rn(u))
        : rng.exp_cold(u);
}

template <bool TRACE>
__global__ void __launch_bounds__(DEEP_BLOCK)
hold_deep_kernel(const DeepArgs d)
{
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ double exp_x[256];                       // ziggurat layer widths (hot table), 32-bit addressed
    for (unsigned i = threadIdx.x; i < 256u; i += blockDim.x) {
        exp_x[i] = zig::zig_exp_x[i];
    }
    __syncthreads();
    uint32_t tab = (uint32_t)__cvta_generic_to_shared(&exp_x[0]);
    asm volatile("" : "+r"(tab));
    const HoldArgs &a = d.h;
    const unsigned lane = threadIdx.x & 31u;
    const uint64_t warps = (uint64_t)gridDim.x * (DEEP_BLOCK / 32);
    const uint64_t gw = (uint64_t)blockIdx.x * (DEEP_BLOCK / 32) + (threadIdx.x >> 5);
    uint4 *const rows = d.rows + gw * d.row_entries;    // this warp's spill area: node i >= 33 at rows[i - 33]
    const uint32_t ticker = (uint32_t)a.workers;
    const uint32_t count0 = ticker + 2u;                // workers + ticker + end event

    for (uint64_t trial = gw; trial < a.num_trials; trial += warps) {
        Sfc64 rng;
        rng.seed(fmix64(a.master_seed, a.first_trial + trial));
        const double mean = a.mean[trial];
        double now = 0.0, sum_wait = 0.0;
        uint32_t pops = 0u, wakes = 0u, ticks = 0u;
        // one variate of look-ahead, as hold_model.cuh; the hot path reads the shared-memory table
        double e_next = std_exponential_shared(rng, tab);

        // cmb_process_start x (workers + 1): START events at t = 0 with keys 1, 2, ... and the end
        // event (key workers + 2) at t = duration.  In index order that is already a heap:
        // equal times with ascending keys, and the one later event last.
        const unsigned long long t_stop = (unsigned long long)__double_as_longlong((double)a.duration);
        uint32_t count = count0;
        unsigned long long rt = 0ull;                   // root = node 0 = key 1 (worker 0, or the ticker if none)
        uint32_t rk = 1u, ri = (0u << 3) | ACT_START;
        if (count0 == 1u) { rt = t_stop; ri = (0xffffu << 3) | 5u; }
        // level 1: node lane + 1
        unsigned long long lt = 0ull;
        uint32_t lk = lane + 2u, li = ((lane + 1u) << 3) | ACT_START;
        if (lane + 2u == count0) { lt = t_stop; li = (0xffffu << 3) | 5u; }
        // deeper levels
        for (uint32_t i = 33u + lane; i < count0; i += 32u) {
            uint4 rec;
            const bool last = i + 1u == count0;
            const unsigned long long t = last ? t_stop : 0ull;
            rec.x = (uint32_t)t;
            rec.y = (uint32_t)(t >> 32);
            rec.z = i + 1u;
            rec.w = last ? ((0xffffu << 3) | 5u) : ((i << 3) | ACT_START);
            rows[i - 33u] = rec;
        }
        uint32_t issued = count0;
        __syncwarp();

        while (count != 0u) {
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
                count = 0u;                             // the end event stops everybody: every pending hold is cancelled
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
            const bool draws = who != ticker;
            const double when = __dadd_rn(now, draws ? __dmul_rn(mean, e_next) : 1.0);
            // the process's next wake-up replaces the root and sinks (dequeue + enqueue fused)
            const unsigned long long mt = (unsigned long long)__double_as_longlong(when);
            const uint32_t mk = ++issued, mi = (who << 3) | ACT_WAKE_TIME;

            // ---- level 1: the lanes' registers
            const bool v1 = lane + 1u < count;
            const Picked p1 = warp_first(v1 ? lt : ~0ull, v1 ? lk : 0xffffffffu);
            if (count == 1u || goes_before(mt, mk, p1.t, p1.key)) {
                rt = mt; rk = mk; ri = mi;
            }
            else {
                rt = p1.t;
                rk = p1.key;
                ri = __shfl_sync(FULL, li, p1.lane);
                const uint32_t n1 = p1.lane + 1u;       // the moving entry now sinks from node n1 (held by lane p1.lane)
                const uint32_t f2 = 32u * n1 + 1u;
                if (f2 >= count) {
                    if (lane == p1.lane) { lt = mt; lk = mk; li = mi; }
                }
                else {
                    // ---- level 2: one coalesced row of 16-byte records
                    const uint32_t c2 = f2 + lane;
                    uint4 r2 = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0u);
                    if (c2 < count) {
                        r2 = rows[c2 - 33u];
                    }
                    const Picked p2 = warp_first(((unsigned long long)r2.y << 32) | r2.x, r2.z);
                    if (goes_before(mt, mk, p2.t, p2.key)) {
                        if (lane == p1.lane) { lt = mt; lk = mk; li = mi; }
                    }
                    else {
                        const uint32_t i2 = __shfl_sync(FULL, r2.w, p2.lane);
                        if (lane == p1.lane) { lt = p2.t; lk = p2.key; li = i2; }
                        const uint32_t n2 = f2 + p2.lane;
                        const uint32_t f3 = 32u * n2 + 1u;
                        uint4 moving = make_uint4((uint32_t)mt, (uint32_t)(mt >> 32), mk, mi);
                        uint32_t at = n2;               // where the moving entry ends up
                        if (f3 < count) {
                            // ---- level 3 (the last one: 33 825 entries at most)
                            const uint32_t c3 = f3 + lane;
                            uint4 r3 = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0u);
                            if (c3 < count) {
                                r3 = rows[c3 - 33u];
                            }
                            const Picked p3 = warp_first(((unsigned long long)r3.y << 32) | r3.x, r3.z);
                            if (!goes_before(mt, mk, p3.t, p3.key)) {
                                if (lane == p3.lane) {
                                    rows[n2 - 33u] = r3;    // the first grandchild moves up
                                }
                                at = f3 + p3.lane;
                            }
                        }
                        if (lane == 0u) {
                            rows[at - 33u] = moving;
                        }
                        __syncwarp();
                    }
                }
            }
            if (draws) {
                e_next = std_exponential_shared(rng, tab);      // refill after use, overlapping the next pop
            }
        }

        if (lane == 0u) {
            if (a.events)    a.events[trial] = pops;
            if (a.objects)   a.objects[trial] = wakes;
            if (a.t_end)     a.t_end[trial] = now;
            if (a.sum_wait)  a.sum_wait[trial] = sum_wait;
            if (a.status)    a.status[trial] = TRIAL_OK;
            if (a.max_queue) a.max_queue[trial] = count0;
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
