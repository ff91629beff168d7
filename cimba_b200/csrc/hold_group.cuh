// hold_group.cuh - the hold model with a G-lane group per trial (G = 8 or 16): 32 / G
// trials advance with every warp instruction.
//
// hold_deep.cuh gives a whole warp to one trial; profiling it (profiles/r01_hold.md) shows
// the kernel issue-bound at ~200 warp instructions per event, most of them the SCALAR part
// of the step (sfc64 + ziggurat, clock, bookkeeping) that all 32 lanes execute redundantly.
// Here the warp is split into 32 / G independent groups.  Each group owns a trial and a
// G-ary heap laid out like hold_deep.cuh's:
//   level 0  the root                     replicated in the group's registers
//   level 1  nodes 1..G                   one node per lane of the group, in registers
//   deeper   node i >= G + 1              global memory (HBM/L2), 16-byte records; the G
//                                         children of a node are one G x 16-byte segment
// so the scalar work is shared by G lanes instead of 32 and one instruction stream serves
// 32 / G events.  A G-ary heap is deeper than a 32-ary one (1002 entries: 4 levels below
// the root at G = 8, 3 at G = 16, 2 at G = 32), which costs more dependent segment reads per
// event but fewer bytes (G x 16 B each); sub-warp REDUX / ballot / shuffle use the group's
// member mask.  Groups of a warp run their trials in lockstep batches (trials of one
// experiment have almost equal event counts); inside a step, groups that sink less deep
// ride along predicated.
//
// Pop order = the reference's (time asc, key asc) whatever the arity,
// so results are bit-identical to hold_model.cuh / hold_deep.cuh and the oracle.
#pragma once

#include "engine.cuh"
#include "hold_deep.cuh"
#include "hold_model.cuh"
#include "rng.cuh"

namespace cimba_b200 {

constexpr int GROUP_BLOCK = 128;

template <int G>
struct GroupPick {
    unsigned lane;              // absolute lane of the first candidate
    unsigned long long t;
    uint32_t key;
};

// first of the group's G candidates under (time asc, key asc); lanes without one pass ~0
template <int G>
__device__ __forceinline__ GroupPick<G> group_first(unsigned gmask, unsigned long long ct, uint32_t ck)
{
    const uint32_t hi = (uint32_t)(ct >> 32), lo = (uint32_t)ct;
    const uint32_t mhi = __reduce_min_sync(gmask, hi);
    unsigned cand = __ballot_sync(gmask, hi == mhi);
    if (__popc(cand) > 1) {                             // ties in the high word (divergent between groups: fine)
        const uint32_t mlo = __reduce_min_sync(gmask, hi == mhi ? lo : 0xffffffffu);
        const bool tie = (hi == mhi) & (lo == mlo);
        const uint32_t mkey = __reduce_min_sync(gmask, tie ? ck : 0xffffffffu);
        cand = __ballot_sync(gmask, tie & (ck == mkey));
    }
    GroupPick<G> p;
    p.lane = __ffs(cand) - 1u;
    p.t = __shfl_sync(gmask, ct, p.lane);
    p.key = __shfl_sync(gmask, ck, p.lane);
    return p;
}

template <int G, bool TRACE>
__global__ void __launch_bounds__(GROUP_BLOCK)
hold_group_kernel(const DeepArgs d)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr uint32_t GROUPS = 32u / G;                // trials per warp
    constexpr uint32_t TOP = G + 1u;                    // nodes held in registers (root + level 1)
    const HoldArgs &a = d.h;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned sub = lane % G;                      // lane within its group
    const unsigned gbase = lane - sub;
    const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << gbase;
    const uint64_t warps = (uint64_t)gridDim.x * (GROUP_BLOCK / 32);
    const uint64_t gw = (uint64_t)blockIdx.x * (GROUP_BLOCK / 32) + (threadIdx.x >> 5);
    const uint64_t slot = gw * GROUPS + lane / G;       // this group's spill area
    uint4 *const rows = d.rows + slot * d.row_entries;  // node i >= TOP at rows[i - TOP]
    const uint32_t ticker = (uint32_t)a.workers;
    const uint32_t count0 = ticker + 2u;                // workers + ticker + end event

    for (uint64_t batch = gw * GROUPS; batch < a.num_trials; batch += warps * GROUPS) {
        const uint64_t trial = batch + lane / G;
        const bool have = trial < a.num_trials;
        Sfc64 rng;
        rng.seed(fmix64(a.master_seed, a.first_trial + (have ? trial : 0u)));
        const double mean = have ? a.mean[trial] : 1.0;
        double now = 0.0, sum_wait = 0.0;
        uint32_t pops = 0u, wakes = 0u, ticks = 0u;
        double e_next = rng.std_exponential_global();

        // initial list in index order (START events at t = 0, keys 1.., the end event last): a valid heap
        const unsigned long long t_stop = (unsigned long long)__double_as_longlong((double)a.duration);
        uint32_t count = have ? count0 : 0u;
        unsigned long long rt = 0ull;
        uint32_t rk = 1u, ri = (0u << 3) | ACT_START;
        unsigned long long lt = 0ull;                   // level 1: node sub + 1
        uint32_t lk = sub + 2u, li = ((sub + 1u) << 3) | ACT_START;
        if (sub + 2u == count0) { lt = t_stop; li = (0xffffu << 3) | 5u; }
        if (have) {
            for (uint32_t i = TOP + sub; i < count0; i += G) {
                uint4 rec;
                const bool last = i + 1u == count0;
                const unsigned long long t = last ? t_stop : 0ull;
                rec.x = (uint32_t)t;
                rec.y = (uint32_t)(t >> 32);
                rec.z = i + 1u;
                rec.w = last ? ((0xffffu << 3) | 5u) : ((i << 3) | ACT_START);
                rows[i - TOP] = rec;
            }
        }
        uint32_t issued = count0;
        __syncwarp();

        while (__any_sync(FULL, count != 0u)) {
            const bool alive = count != 0u;
            bool sinking = false;
            unsigned long long mt = 0ull;
            uint32_t mk = 0u, mi = 0u;
            bool draws = false;
            if (alive) {
                now = __longlong_as_double((long long)rt);
                if (TRACE) {
                    if (sub == 0u && pops < a.trace_cap) {
                        a.trace_key[trial * a.trace_cap + pops] = rk;
                        a.trace_time[trial * a.trace_cap + pops] = now;
                    }
                }
                pops++;
                const uint32_t act = ri & 7u, who = ri >> 3;
                if (act == 5u) {
                    count = 0u;                         // the end event: every pending hold is cancelled
                }
                else {
                    if (act == ACT_WAKE_TIME) {
                        if (who == ticker) {
                            ticks++;
                        }
                        else {
                            wakes++;
                            sum_wait = __dadd_rn(sum_wait, now);
                        }
                    }
                    draws = who != ticker;
                    const double when = __dadd_rn(now, draws ? __dmul_rn(mean, e_next) : 1.0);
                    mt = (unsigned long long)__double_as_longlong(when);
                    mk = ++issued;
                    mi = (who << 3) | ACT_WAKE_TIME;
                    sinking = true;                     // the next wake-up replaces the root and sinks
                }
            }

            // ---- level 1: the group's registers
            const bool v1 = sinking & (sub + 1u < count);
            const GroupPick<G> p1 = group_first<G>(gmask, v1 ? lt : ~0ull, v1 ? lk : 0xffffffffu);
            uint32_t node = 0u;                         // where the moving entry currently sits
            if (sinking) {
                if (count == 1u || goes_before(mt, mk, p1.t, p1.key)) {
                    rt = mt; rk = mk; ri = mi;
                    sinking = false;
                }
            }
            const uint32_t i1 = __shfl_sync(gmask, li, p1.lane);
            if (sinking) {
                rt = p1.t; rk = p1.key; ri = i1;
                node = p1.lane - gbase + 1u;
            }
            const unsigned holder = p1.lane;            // the lane whose registers are node `node` while node <= G

            // ---- deeper levels: one G x 16-byte segment per level
            for (;;) {
                const uint32_t first = G * node + 1u;
                const bool has_kids = sinking & (first < count);
                if (sinking & !has_kids) {              // a leaf: the moving entry stays here
                    if (node <= G) {
                        if (lane == holder) { lt = mt; lk = mk; li = mi; }
                    }
                    else if (sub == 0u) {
                        rows[node - TOP] = make_uint4((uint32_t)mt, (uint32_t)(mt >> 32), mk, mi);
                    }
                    sinking = false;
                }
                if (!__any_sync(FULL, has_kids)) {
                    break;
                }
                const uint32_t c = first + sub;
                uint4 r = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0u);
                if (has_kids && c < count) {
                    r = rows[c - TOP];
                }
                const GroupPick<G> p = group_first<G>(gmask, ((unsigned long long)r.y << 32) | r.x, r.z);
                const uint32_t pinfo = __shfl_sync(gmask, r.w, p.lane);
                if (has_kids) {
                    if (goes_before(mt, mk, p.t, p.key)) {
                        if (node <= G) {
                            if (lane == holder) { lt = mt; lk = mk; li = mi; }
                        }
                        else if (sub == 0u) {
                            rows[node - TOP] = make_uint4((uint32_t)mt, (uint32_t)(mt >> 32), mk, mi);
                        }
                        sinking = false;
                    }
                    else {                              // the first child moves up into `node`
                        if (node <= G) {
                            if (lane == holder) { lt = p.t; lk = p.key; li = pinfo; }
                        }
                        else if (lane == p.lane) {
                            rows[node - TOP] = r;
                        }
                        node = first + (p.lane - gbase);
                    }
                }
                __syncwarp();
            }
            __syncwarp();
            if (draws) {
                e_next = rng.std_exponential_global();
            }
        }

        if (have && sub == 0u) {
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
