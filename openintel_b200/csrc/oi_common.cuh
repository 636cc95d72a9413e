// oi_common.cuh — shared device helpers: 64-bit ranking keys (SPEC §1), block-wide bitonic
// sort, the block candidate buffer used by every fused top-k, and the PTX wrappers for
// mbarrier / bulk async copy (TMA's 1-D path) on sm_100a.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef unsigned long long u64;

#define OI_NO_DOC_U32 0xFFFFFFFFu
#define OI_SEL_CAP 2048  // candidate-buffer capacity (keys) of one CTA; max k = OI_SEL_CAP / 2

// ---------------------------------------------------------------------------------------------
// SPEC §1 keys.  key 0 is the "empty" sentinel: no finite score maps to ord == 0.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t oi_ord(float s) {
  s = s + 0.0f;  // canon: -0.0 -> +0.0
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(s);
#else
  uint32_t u;
  memcpy(&u, &s, 4);
#endif
  return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__host__ __device__ __forceinline__ u64 oi_make_key(float s, uint32_t doc) {
  return ((u64)oi_ord(s) << 32) | (u64)(0xFFFFFFFFu - doc);
}
__host__ __device__ __forceinline__ uint32_t oi_key_doc(u64 key) { return 0xFFFFFFFFu - (uint32_t)key; }
__host__ __device__ __forceinline__ float oi_key_score(u64 key) {
  uint32_t u = (uint32_t)(key >> 32);
  u ^= (u >> 31) ? 0x80000000u : 0xFFFFFFFFu;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

// ---------------------------------------------------------------------------------------------
// named barrier over the first `nthreads` threads' warps (id 0 with all threads == __syncthreads)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void oi_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ uint32_t oi_next_pow2(uint32_t x) {
  return x <= 1 ? 1u : 1u << (32 - __clz(x - 1));
}

// Bitonic sort of buf[0..n) (n a power of two), DESCENDING, by `nthreads` threads that all call
// this with tid in [0, nthreads) and synchronise through named barrier `bar`.
__device__ __forceinline__ void oi_bitonic_desc(u64 *buf, uint32_t n, int tid, int nthreads, int bar) {
  for (uint32_t k = 2; k <= n; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t i = tid; i < (n >> 1); i += nthreads) {
        uint32_t l = ((i & ~(j - 1)) << 1) | (i & (j - 1));
        uint32_t r = l | j;
        u64 a = buf[l], b = buf[r];
        bool desc = (l & k) == 0;
        if ((a < b) == desc) { buf[l] = b; buf[r] = a; }
      }
      oi_bar_sync(bar, nthreads);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Block candidate buffer ("filter + buffer" selection).
//   buf[OI_SEL_CAP] keys, *cnt entries used, *thr = a key such that at least k candidates with
//   key >= *thr have been seen (0 = no bound yet).  Producers test `key > *thr` (or >=, both
//   valid) and push; between synchronisation points the caller guarantees that at most
//   OI_SEL_CAP - *cnt pushes happen.  compact() keeps the best k.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void oi_sel_push(u64 *buf, uint32_t *cnt, u64 key) {
  uint32_t p = atomicAdd(cnt, 1u);
  if (p < OI_SEL_CAP) buf[p] = key;  // the bound is the caller's invariant; never write outside
}

// All `nthreads` threads call (after a barrier that made the pushes visible).  On return (with
// a trailing barrier done) buf[0..*cnt) is sorted descending, *cnt <= k, *thr updated.
__device__ __forceinline__ void oi_sel_compact(u64 *buf, uint32_t *cnt, u64 *thr, uint32_t k, int tid,
                                               int nthreads, int bar) {
  uint32_t c = min(*cnt, (uint32_t)OI_SEL_CAP);
  uint32_t n = oi_next_pow2(c);
  for (uint32_t i = c + tid; i < n; i += nthreads) buf[i] = 0ull;
  oi_bar_sync(bar, nthreads);
  oi_bitonic_desc(buf, n, tid, nthreads, bar);
  if (tid == 0) {
    uint32_t keep = min(c, k);
    *cnt = keep;
    if (keep == k && buf[k - 1] > *thr) *thr = buf[k - 1];
  }
  oi_bar_sync(bar, nthreads);
}

// One WARP keeps the best k of buf[0..*cnt) (sorted descending) and raises *thr when k are held; no CTA barrier
// is involved, so several warps can compact different buffers at the same time.  cap = buffer capacity.
__device__ __forceinline__ void oi_warp_sel_compact(u64 *buf, uint32_t *cnt, u64 *thr, uint32_t cap, uint32_t k, int lane) {
  const uint32_t c = min(*cnt, cap);
  const uint32_t n = oi_next_pow2(c);
  __syncwarp();
  for (uint32_t i = c + lane; i < n; i += 32) buf[i] = 0ull;
  __syncwarp();
  for (uint32_t kk = 2; kk <= n; kk <<= 1) {
    for (uint32_t j = kk >> 1; j > 0; j >>= 1) {
      for (uint32_t i = lane; i < (n >> 1); i += 32) {
        const uint32_t l = ((i & ~(j - 1)) << 1) | (i & (j - 1));
        const uint32_t r = l | j;
        const u64 a = buf[l], b = buf[r];
        const bool desc = (l & kk) == 0;
        if ((a < b) == desc) { buf[l] = b; buf[r] = a; }
      }
      __syncwarp();
    }
  }
  if (lane == 0) {
    const uint32_t keep = min(c, k);
    *cnt = keep;
    if (keep == k && buf[k - 1] > *thr) *thr = buf[k - 1];
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// mbarrier + bulk async copy (cp.async.bulk, SASS UBLKCP) wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t oi_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void oi_mbar_init(void *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(oi_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void oi_mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void oi_mbar_expect_tx(void *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(oi_smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void oi_mbar_arrive(void *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(oi_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void oi_mbar_wait(void *bar, uint32_t parity) {
  uint32_t a = oi_smem_u32(bar);
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(a),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy, completion reported to `bar` as transaction bytes.
// bytes % 16 == 0, both addresses 16 B aligned.  L2 policy: evict-first (pure streaming).
__device__ __forceinline__ void oi_bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, void *bar,
                                            u64 policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(oi_smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(oi_smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ u64 oi_policy_evict_first() {
  u64 p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// 128-bit streaming load that does not allocate in L1
__device__ __forceinline__ uint4 oi_ldg_stream(const uint4 *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
