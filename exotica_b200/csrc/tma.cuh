// tma.cuh -- device-side TMA store primitives (sm_100a): tensor stores of shared-memory tiles, bulk-group completion, the
// generic->async proxy fence and an L2 evict-first policy (Phi / J are written once and never re-read by the producing kernel).
#pragma once
#include <stdint.h>

#include "kernels.cuh"

namespace exo
{
namespace tma
{
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
// box of the tensor map's size from shared memory (dense [box_rows][box_inner]) to coordinates (c_inner, c_row); elements
// outside the tensor's extents are clipped by the hardware (ragged last tile)
__device__ __forceinline__ void store_2d(const TmaDesc* map, const void* smem, int c_inner, int c_row, uint64_t policy)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;" ::"l"(map), "r"(smem_u32(smem)),
                 "r"(c_inner), "r"(c_row), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void commit()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// all bulk groups of this thread have finished READING shared memory (the buffer may be reused / the CTA may exit)
__device__ __forceinline__ void wait_read_all()
{
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// makes this thread's st.shared visible to the async proxy (the TMA unit); every staging thread runs it before the barrier
__device__ __forceinline__ void fence_staged()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
}  // namespace tma
}  // namespace exo
