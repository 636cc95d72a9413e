// oi_tcgen05.cuh — thin PTX wrappers for the sm_100a tensor-core path: TMEM allocation,
// tcgen05.mma (A from TMEM, B from shared memory), tcgen05.ld/st, tcgen05.commit, and the
// 2-D TMA tile load (cp.async.bulk.tensor).  SASS: UTCHMMA / LDTM / STTM / UTMALDG.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "oi_common.cuh"

// ---- TMEM ----------------------------------------------------------------------------------------
// whole-warp calls (.sync.aligned)
__device__ __forceinline__ void oi_tmem_alloc(uint32_t *smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oi_smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void oi_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void oi_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void oi_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void oi_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void oi_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive 32-bit columns: thread t of the warp gets lane (taddr.lane + t)
__device__ __forceinline__ void oi_tmem_ld32(uint32_t taddr, uint32_t *r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void oi_tmem_st32(uint32_t taddr, const uint32_t *r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

// ---- descriptors ---------------------------------------------------------------------------------
// kind::f16 instruction descriptor: D f32, A/B bf16, both K-major, dense.
//   bits [4,6) c_format = 1 (F32); [7,10) a_format = 1 (BF16); [10,13) b_format = 1 (BF16);
//   bit 15 / 16 a_major / b_major = 0 (K); [17,23) N >> 3; [24,29) M >> 4.
__host__ __device__ constexpr uint32_t oi_umma_idesc_bf16(uint32_t M, uint32_t N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// Shared-memory matrix descriptor of a K-major tile whose rows are 128 B (64 bf16) with the
// 128-byte swizzle (the layout a SWIZZLE_128B TMA box lands in): 8-row groups 1024 B apart
// (SBO = 64 x 16 B), LBO = 1 (unused for swizzled K-major), version 1 (sm_100), layout type 2.
__device__ __forceinline__ uint64_t oi_umma_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)1u << 16;
  d |= (uint64_t)64u << 32;
  d |= (uint64_t)1u << 46;
  d |= (uint64_t)2u << 61;
  return d;
}

// ---- MMA -----------------------------------------------------------------------------------------
// D[tmem] (+)= A[tmem] * B[smem]^T : one thread issues for the CTA
__device__ __forceinline__ void oi_umma_ts_bf16(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the mbarrier gets one arrival when every MMA issued so far by this thread has completed
__device__ __forceinline__ void oi_umma_commit(void *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(oi_smem_u32(bar))
               : "memory");
}

// ---- TMA -----------------------------------------------------------------------------------------
__device__ __forceinline__ void oi_tma_prefetch_desc(const CUtensorMap *m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
// box at (c0 = element index in the row, c1 = row) -> shared memory, completion on `bar`
__device__ __forceinline__ void oi_tma_load_2d(void *smem_dst, const CUtensorMap *m, int32_t c0, int32_t c1, void *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          oi_smem_u32(smem_dst)),
      "l"(m), "r"(oi_smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// 3-D box at (c0, c1, c2): used with the (64 elements, rows, k-blocks) view of a row-major matrix
__device__ __forceinline__ void oi_tma_load_3d(void *smem_dst, const CUtensorMap *m, int32_t c0, int32_t c1, int32_t c2, void *bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          oi_smem_u32(smem_dst)),
      "l"(m), "r"(oi_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ---- CTA pairs (cta_group::2): two CTAs of a cluster on the two SMs of a TPC issue ONE MMA of M = 256 ---------------
// Each CTA holds its 128 rows of A (tensor memory) and of D, and HALF of B's rows (N / 2) in its shared memory; the
// leader CTA (cluster rank 0) issues the instruction for both.
__device__ __forceinline__ uint32_t oi_cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void oi_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// whole-warp calls, executed by the same warp of BOTH CTAs of the pair
__device__ __forceinline__ void oi_tmem_alloc2(uint32_t *smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oi_smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void oi_tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem, both CTAs] (+)= A[tmem, both CTAs] * B[smem halves of both CTAs]^T : one thread of the leader CTA issues
__device__ __forceinline__ void oi_umma2_ts_bf16(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one arrival on the mbarrier at this shared-memory offset in EVERY CTA of `cta_mask` when the MMAs issued so far complete
__device__ __forceinline__ void oi_umma2_commit_mc(void *bar, uint32_t cta_mask) {
  asm volatile(
      "{\n\t"
      ".reg .b16 m;\n\t"
      "cvt.u16.u32 m, %1;\n\t"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t"
      "}" ::"r"(oi_smem_u32(bar)),
      "r"(cta_mask)
      : "memory");
}
// 3-D TMA box into this CTA's shared memory; the transaction bytes are reported to the mbarrier at the same offset in
// the LEADER CTA of the pair (the cluster-window address with the peer bit cleared)
__device__ __forceinline__ void oi_tma_load_3d_pair(void *smem_dst, const CUtensorMap *m, int32_t c0, int32_t c1, int32_t c2, void *bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          oi_smem_u32(smem_dst)),
      "l"(m), "r"(oi_smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// arrive on the mbarrier at this offset in CTA `cta` of the cluster.  Default semantics (release at CTA scope), as
// CUTLASS's ClusterBarrier::arrive: what the barrier hands over here is tensor memory, ordered by the tcgen05 fences
// on both sides -- a `.release.cluster` arrive costs a device-wide MEMBAR per call (37 % of the pair kernel's stall
// samples: profiles/r02_gemm_ab.md).
__device__ __forceinline__ void oi_mbar_arrive_cluster(void *bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(oi_smem_u32(bar)),
      "r"(cta)
      : "memory");
}
// wait on a barrier the peer CTA arrives on (see oi_mbar_arrive_cluster: tensor memory only, plain wait)
__device__ __forceinline__ void oi_mbar_wait_cluster(void *bar, uint32_t parity) { oi_mbar_wait(bar, parity); }
