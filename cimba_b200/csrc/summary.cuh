// summary.cuh - cmb_datasummary arithmetic (host + device) and the on-device
// reduction of per-trial results.
//
// Reference: src/cmb_datasummary.c:144-166 (add: Pebay's single-sample update in
// Meng's evaluation order) and :93-131 (merge: Pebay's pairwise formula).  The
// benchmark folds the per-trial averages with a serial host loop
// (benchmark/MM1_multi.c:143-148); here each GPU folds its shard with a fixed,
// deterministic tree of the same two operations, and ranks are merged in rank
// order after an all-gather.  A merged tree differs from
// the serial fold by rounding only (<< 1e-9 relative).
//
// Compiled with -fmad=false: the expressions below keep the reference's
// operation order and must not be contracted.
#pragma once

#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

namespace cimba_b200 {

struct SummaryAcc {
    uint64_t count;
    double min, max, m1, m2, m3, m4;
};

__host__ __device__ inline SummaryAcc summary_empty()
{
    return SummaryAcc{0u, DBL_MAX, -DBL_MAX, 0.0, 0.0, 0.0, 0.0};
}

// cmb_datasummary_add, src/cmb_datasummary.c:144-166
__host__ __device__ inline void summary_add(SummaryAcc &s, double y)
{
    s.max = (y > s.max) ? y : s.max;
    s.min = (y < s.min) ? y : s.min;

    const double d = y - s.m1;
    const double d_2 = d * d;
    const double d_3 = d * d_2;
    const double n = (double)(++s.count);
    const double d_n = d / n;
    const double d_n_2 = d_n * d_n;
    const double d_n_3 = d_n_2 * d_n;

    s.m1 += d_n;
    s.m2 += d * (d - d_n);
    s.m3 += d * (d_2 - d_n_2) - 3.0 * d_n * s.m2;
    s.m4 += d * (d_3 - d_n_3) - 6.0 * d_n_2 * s.m2 - 4.0 * d_n * s.m3;
}

// cmb_datasummary_merge, src/cmb_datasummary.c:93-131.  An empty side is passed
// through unchanged (the reference formula would divide 0/0 for two empties).
__host__ __device__ inline SummaryAcc summary_merge(const SummaryAcc &a, const SummaryAcc &b)
{
    if (b.count == 0u) {
        return a;
    }
    if (a.count == 0u) {
        return b;
    }
    SummaryAcc c;
    c.count = a.count + b.count;
    c.min = (a.min < b.min) ? a.min : b.min;
    c.max = (a.max > b.max) ? a.max : b.max;

    const double n1 = (double)a.count;
    const double n2 = (double)b.count;
    const double n = (double)c.count;
    const double d21 = b.m1 - a.m1;
    const double d21_n = d21 / n;
    const double d21_n_2 = d21_n * d21_n;
    const double d21_n_3 = d21_n * d21_n_2;

    c.m1 = a.m1 + n2 * d21_n;
    c.m2 = a.m2 + b.m2 + n1 * n2 * d21 * d21_n;
    c.m3 = a.m3 + b.m3
         + n1 * n2 * (n1 - n2) * d21 * d21_n_2
         + 3.0 * (n1 * b.m2 - n2 * a.m2) * d21_n;
    c.m4 = a.m4 + b.m4
         + n1 * n2 * (n1 * n1 - n1 * n2 + n2 * n2) * d21 * d21_n_3
         + 6.0 * (n1 * n1 * b.m2 + n2 * n2 * a.m2) * d21_n_2
         + 4.0 * (n1 * b.m3 - n2 * a.m3) * d21_n;
    return c;
}

constexpr int SUMMARY_BLOCK = 256;

// One CTA: thread t adds trials t, t+256, ... (coalesced reads), then a fixed
// halving tree of merges in shared memory.  out = {count,min,max,m1,m2,m3,m4,0}.
__global__ void __launch_bounds__(SUMMARY_BLOCK)
summarize_kernel(const double *__restrict__ sum_wait, const uint64_t *__restrict__ objects,
                 uint64_t n, double *__restrict__ out)
{
    __shared__ SummaryAcc part[SUMMARY_BLOCK];
    SummaryAcc acc = summary_empty();
    for (uint64_t i = threadIdx.x; i < n; i += SUMMARY_BLOCK) {
        // avg_tsys = sum_wait / (double)obj_cnt, benchmark/MM1_multi.c:146
        summary_add(acc, sum_wait[i] / (double)objects[i]);
    }
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int s = SUMMARY_BLOCK / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            part[threadIdx.x] = summary_merge(part[threadIdx.x], part[threadIdx.x + s]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const SummaryAcc r = part[0];
        out[0] = (double)r.count;
        out[1] = r.min;
        out[2] = r.max;
        out[3] = r.m1;
        out[4] = r.m2;
        out[5] = r.m3;
        out[6] = r.m4;
        out[7] = 0.0;
    }
}

}  // namespace cimba_b200
