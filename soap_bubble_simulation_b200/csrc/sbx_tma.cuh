// sbx_tma.cuh -- mbarrier / TMA bulk-copy primitives shared by the ring kernels (inline PTX for sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "sbx_internal.cuh"

namespace sbx {
namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Hand a ring stage back to the producer AFTER this warp's shared-memory reads of it have returned.
// A plain `ld.shared ...; mbarrier.arrive` pair is not enough: nothing makes the arrive wait for the loads' data
// (SASS: SYNCS.ARRIVE carries no scoreboard wait on the preceding LDS), so under load-store-unit back-pressure the
// arrive -- and the TMA overwrite it allows -- can overtake a load that is still queued, and the warp then reads the
// NEXT contents of the slot (seen as one wrong row segment per ~1e6 warp-rows at 8192 x 16384).
// The releasing lane therefore first STORES an OR of the loaded registers (`bits`) to a per-warp sink word in shared
// memory and arrives after that: the store cannot issue before every load that produced `bits` has written its
// registers (a true register dependency, for all lanes, since a warp's LDS is one instruction), and the arrive has
// release semantics, which orders it after the thread's earlier store.  No assumption about the block shape and
// nothing a compiler stage can fold away.  (SBX_ARRIVE_SINK=0 builds the round-1 form, which carried the dependency
// in the barrier ADDRESS through a run-time zero, blockDim.y - 1: same measured behaviour, 1-D blocks only.)
#ifndef SBX_ARRIVE_SINK
#define SBX_ARRIVE_SINK 1
#endif
__device__ __forceinline__ uint32_t dep_bits(float v) { return __float_as_uint(v); }
__device__ __forceinline__ uint32_t dep_bits(double v) { return (uint32_t)__double2hiint(v) | (uint32_t)__double2loint(v); }
__device__ __forceinline__ void mbar_arrive_after(uint64_t *bar, uint32_t bits)
{
#if SBX_ARRIVE_SINK
    __shared__ uint32_t s_sink[32];  // one word per warp of the block
    asm volatile(
        "st.volatile.shared.u32 [%1], %2;\n\t"
        "mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)),
        "r"(smem_u32(&s_sink[threadIdx.x >> 5])), "r"(bits)
        : "memory");
#else
    const uint32_t off = bits & (blockDim.y - 1u);
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar) + off) : "memory");
#endif
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared; completion is counted in bytes on an mbarrier
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// same copy with an L2 eviction-priority policy (createpolicy), e.g. evict_last for rows a neighbouring band re-reads
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_bulk_g2s_hint(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar,
                                                  uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void bar_consumers(int nthreads)
{
    asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

// ---- peer-memory halo (PeerArgs): ordering between latitude neighbours, shared by the ring kernels
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// wait until a neighbour has finished the previous step (bounded: a dead neighbour must not hang the GPU)
__device__ __forceinline__ void wait_neighbour(const unsigned long long *flag, unsigned long long want, unsigned int *status)
{
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(flag) < want) {
        if (globaltimer_ns() - t0 > 5000000000ull) {  // 5 s
            atomicOr(status, 1u);
            break;
        }
        __nanosleep(64);
    }
    asm volatile("fence.proxy.async;" ::: "memory");  // the rows the neighbour pushed are read by TMA next
}

// When every CTA of my first (last) band is done -- the neighbour's row consumed, my boundary row written and fenced --
// publish that this step is finished here; the up (down) neighbour polls the word over NVLink before it reads my row
// and before it overwrites the row I have just read.  Called by all consumer threads of a CTA at the end of its band.
__device__ __forceinline__ void peer_signal_done(const PeerArgs &p, bool sig_up, bool sig_dn, int consumers, int tid)
{
    if (!(sig_up || sig_dn)) return;
#ifdef SBX_PEER_DEBUG
    if (p.debug & 4u) return;
    if (!(p.debug & 8u))
#endif
    __threadfence_system();
    bar_consumers(consumers);
    if (tid == 0) {
        if (sig_up && atomicAdd(p.done_first, 1u) == gridDim.x - 1) {
            *p.done_first = 0u;
            st_release_sys(p.my_first_done, p.epoch);
        }
        if (sig_dn && atomicAdd(p.done_last, 1u) == gridDim.x - 1) {
            *p.done_last = 0u;
            st_release_sys(p.my_last_done, p.epoch);
        }
    }
}


}  // namespace
}  // namespace sbx
