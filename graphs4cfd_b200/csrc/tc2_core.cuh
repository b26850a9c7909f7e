// tc2_core.cuh — second-generation tensor-core primitives (sm_100a): CTA pairs (cta_group::2),
// A operands resident in TMEM (".ts" form of tcgen05.mma), smem -> TMEM copies (tcgen05.cp) used as a
// hardware transposer for coalesced row loads, tcgen05.st, cluster-scope mbarrier traffic.
//
// Conventions shared by every kernel built on this header
//  * "image" = 128 B-pitch rows, SWIZZLE_128B (16-byte chunk c of row r stored at chunk c ^ (r & 7)),
//    8-row groups 1024 B apart, base 1024-byte aligned: the canonical K-major UMMA operand layout.
//    An image of R rows is R*128 bytes.  The same byte format serves fp16 MMA operands (64 k per row),
//    and fp32 tiles (32 columns per row) that tcgen05.cp moves into TMEM accumulator columns.
//  * TMEM address = (lane << 16) | column; warp w may touch lanes 32*(w%4) .. +31 only.
//  * A operand in TMEM: lane = row, column c holds k = 2c (low half) and k = 2c+1 (high half).
//  * every mbarrier wait is bounded: a wait that spins for seconds traps instead of hanging the GPU.
#pragma once
#include "tc_core.cuh"

namespace g4c {
namespace tc2 {

using tc::fence_barrier_init;
using tc::fence_proxy_async;
using tc::make_desc_sw128;
using tc::mbar_arrive;
using tc::mbar_arrive_expect_tx;
using tc::mbar_init;
using tc::smem_u32;
using tc::split2;
using tc::tc_fence_after;
using tc::tc_fence_before;
using tc::tmem_ld32;
using tc::bulk_g2s;

constexpr uint32_t kWatchdogSpins = 1u << 24;

// kind::f16 instruction descriptor: D=f32, A=B=f16, both K-major
__host__ __device__ constexpr uint32_t idesc_f16(uint32_t M, uint32_t N) {
    return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ------------------------------------------------------------------ cluster
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same variable in CTA `cta` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
    return r;
}
// arrive on an mbarrier of another CTA of the cluster (address from mapa).  Default semantics (.release.cta), the
// form CUTLASS uses for 2-CTA pipelines: the data handed over lives in TMEM and is ordered by
// tcgen05.wait::st + tcgen05.fence::before_thread_sync, so no MEMBAR.ALL.GPU (which the .release.cluster form
// emits, and which would also wait for every global store of the thread) is needed.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) { mbar_arrive_remote(cluster_addr); }

// One elected lane of a converged warp.  Code under `if (elect_one())` is compiled for the uniform datapath: the MMA issue
// loop becomes three back-to-back UTCHMMA per K step with uniform-register descriptor arithmetic, where `if (lane == 0)`
// costs ~100 vector instructions and a lane loop per K step (the issuing thread then needs ~110 cycles per MMA).
__device__ __forceinline__ bool elect_one() {
    uint32_t p;
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}\n" : "=r"(p));
    return p != 0;
}

// ------------------------------------------------------------------ bounded waits
// plain (cta-scope acquire) try_wait: the data handed over through these barriers lives in TMEM or was written by
// the async proxy, and is ordered by tcgen05 fences / complete_tx; a .acquire.cluster wait would add a CCTL.IVALL
// (L1 invalidation) to every successful wait.
template <bool kClusterScope>
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
template <bool kClusterScope = false>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try<kClusterScope>(bar, parity)) {
        if (++spins > kWatchdogSpins) __trap();
    }
}

// Waiting without taking issue slots from the working warps of the same scheduler: try_wait with a suspend-time
// hint compiles to TRYWAIT + NANOSLEEP.SYNCS (the warp sleeps until mbarrier activity or the hint expires).  A plain
// try_wait loop re-issues every ~20 cycles, which with several waiting warps per scheduler starves the others.
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0, ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");
        if (!ok && ++spins > kWatchdogSpins) __trap();
    } while (!ok);
}

// ------------------------------------------------------------------ explicit shared-space accesses by 32-bit address
// In a cluster kernel a dereference of a generic pointer into shared memory costs an S2R of the CTA's window id
// plus address arithmetic at every access; hot loops use these instead (address = smem_u32(base) + offset, once).
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds_f1(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f1(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void mbar_arrive_a(uint32_t addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_sleep_a(uint32_t addr, uint32_t parity) {
    uint32_t spins = 0, ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok) : "r"(addr), "r"(parity), "r"(20000u) : "memory");
        if (!ok && ++spins > kWatchdogSpins) __trap();
    } while (!ok);
}
template <int CG>
__device__ __forceinline__ void umma_commit_a(uint32_t bar_addr, uint16_t mask = 3) {
    if (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar_addr),
                     "h"(mask)
                     : "memory");
}

// ------------------------------------------------------------------ TMEM management
template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_holder, uint32_t ncols) {
    if (CG == 1)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_holder)), "r"(ncols) : "memory");
    else
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_holder)), "r"(ncols) : "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_relinquish() {
    if (CG == 1) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    else asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ------------------------------------------------------------------ MMA with the A operand in TMEM
// D[tmem] (+)= A[tmem] * B[smem]^T ; one K=16 step.  CG=2: issued by the leader CTA for the pair
// (M = 256: each CTA contributes its own 128 TMEM lanes of A and D and its own half of B's rows).
template <int CG>
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    if (CG == 1)
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
            "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
            "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
}
// arrive on `bar` (same smem offset in every CTA of `mask` for CG=2) when all prior tcgen05 async ops
// (mma, cp) issued by this thread have completed
template <int CG>
__device__ __forceinline__ void umma_commit(uint64_t* bar, uint16_t mask = 3) {
    if (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                     "h"(mask)
                     : "memory");
}
// smem image -> TMEM: 128 rows x 32 bytes (8 TMEM columns) starting at the descriptor's start address
template <int CG>
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t s_desc) {
    if (CG == 1) asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(s_desc) : "memory");
    else asm volatile("tcgen05.cp.cta_group::2.128x256b [%0], %1;" ::"r"(taddr), "l"(s_desc) : "memory");
}

// ------------------------------------------------------------------ registers -> TMEM
// 32 consecutive 32-bit columns of this thread's lane
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 fp32 columns straight into a float array slice (waits for completion)
__device__ __forceinline__ void tmem_ld32f(uint32_t taddr, float* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
          "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]),
          "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]),
          "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
        : "r"(taddr)
        : "memory");
}

// 16 fp32 columns, no implicit wait (caller batches loads, then tmem_wait_ld())
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
          "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16f(uint32_t taddr, float* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
          "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
        : "r"(taddr)
        : "memory");
}

// byte offset of 16-byte chunk `c` (0..7) of row r in an image
__device__ __forceinline__ uint32_t img_off(int r, int c) { return (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4); }

}  // namespace tc2
}  // namespace g4c
