// tma_bulk.cuh - the TMA bulk-copy engine (cp.async.bulk, non-tensor form) and its mbarrier completion, as the few
// primitives the simulation kernels need: a contiguous run of bytes global -> shared::cta signalled on an mbarrier
// (stage-in / refill), shared::cta -> global in a bulk group (stage-out / spill), and the proxy fences that order the
// asynchronous proxy against ordinary loads and stores.  sm_90+ PTX; compiled here for sm_100a only.
//
// Rules the callers follow (PTX ISA, "Asynchronous copy" and "mbarrier"):
//   * source, destination and size are multiples of 16 bytes;
//   * one thread arms the barrier with the number of bytes in flight (arrive.expect_tx) and issues the copies; every
//     consumer waits on the barrier's phase parity, which flips once per completed (arrival count + bytes) round;
//   * data written with ordinary stores and then read by a bulk copy (or the reverse through shared memory) needs a
//     fence.proxy.async in between - the copy engine is a different memory proxy.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace cimba_b200 {
namespace tma {

__device__ __forceinline__ uint32_t smem_addr(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void barrier_init(uint64_t *bar, uint32_t arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(arrivals) : "memory");
}

// make freshly initialised barriers visible to the copy engine
__device__ __forceinline__ void barrier_init_fence()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// one arrival that also announces `bytes` of asynchronous traffic to come
__device__ __forceinline__ void barrier_expect(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr(bar)), "r"(bytes) : "memory");
}

// spin until the phase with the given parity has completed
__device__ __forceinline__ void barrier_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(smem_addr(bar)), "r"(parity) : "memory");
}

// global -> this CTA's shared memory; completion (bytes) is counted on `bar`
__device__ __forceinline__ void load(void *dst_smem, const void *src_global, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_addr(dst_smem)), "l"(src_global), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

// this CTA's shared memory -> global, as part of the thread's current bulk group
__device__ __forceinline__ void store(void *dst_global, const void *src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(dst_global), "r"(smem_addr(src_smem)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void store_commit()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// all of this thread's bulk stores have READ their shared-memory source (it may be overwritten) ...
__device__ __forceinline__ void store_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ... and have completed (their global destination is written)
__device__ __forceinline__ void store_wait()
{
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ordinary stores to shared memory before a bulk store reads them
__device__ __forceinline__ void fence_smem_to_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ordinary stores to global memory before a bulk load reads them
__device__ __forceinline__ void fence_global_to_async()
{
    asm volatile("fence.proxy.async.global;" ::: "memory");
}

}  // namespace tma
}  // namespace cimba_b200
