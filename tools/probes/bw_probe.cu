// bw_probe.cu — read-bandwidth ceiling probe for the cosine scan (not part of the product).
// Streams a 1.536 GB buffer (the config-2 matrix size) with several load shapes and prints GB/s:
//   static  : contiguous range per CTA (what cosine_scan_ldg_kernel v0 does)
//   stride  : grid-stride at warp granularity (all CTAs interleaved)
//   dynamic : 48 KB tiles handed out by an atomic counter
//   bulk    : cp.async.bulk into a shared-memory ring, dynamic tiles
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bw_probe bw_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ uint4 ldg_stream(const uint4 *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
template <int U>
__global__ void __launch_bounds__(256) k_static(const uint4 *m, u64 nvec, uint32_t *out) {
  const u64 per = (nvec + gridDim.x - 1) / gridDim.x;
  const u64 b = per * blockIdx.x, e = min(nvec, b + per);
  uint32_t acc = 0;
  for (u64 i = b + threadIdx.x; i < e; i += 256 * U) {
    uint4 v[U];
#pragma unroll
    for (int j = 0; j < U; ++j) v[j] = (i + 256ull * j < e) ? ldg_stream(m + i + 256ull * j) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int j = 0; j < U; ++j) acc += v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
  }
  if (acc == 0x12345678u) out[0] = acc;
}
template <int U>
__global__ void __launch_bounds__(256) k_stride(const uint4 *m, u64 nvec, uint32_t *out) {
  uint32_t acc = 0;
  const u64 step = (u64)gridDim.x * 256 * U;
  for (u64 i = (u64)blockIdx.x * 256 * U + threadIdx.x; i < nvec; i += step) {
    uint4 v[U];
#pragma unroll
    for (int j = 0; j < U; ++j) v[j] = (i + 256ull * j < nvec) ? ldg_stream(m + i + 256ull * j) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int j = 0; j < U; ++j) acc += v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
  }
  if (acc == 0x12345678u) out[0] = acc;
}
// dynamic tiles of TV vectors; warp w of the CTA takes rows of 96 vectors (1536 B) like the scan
template <int U>
__global__ void __launch_bounds__(256) k_dynamic(const uint4 *m, u64 nvec, uint32_t *ctr, uint32_t tile_vecs, uint32_t *out) {
  __shared__ uint32_t s_tile;
  uint32_t acc = 0;
  const u64 n_tiles = (nvec + tile_vecs - 1) / tile_vecs;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_tile = atomicAdd(ctr, 1u);
    __syncthreads();
    const u64 t = s_tile;
    if (t >= n_tiles) break;
    const u64 b = t * tile_vecs, e = min(nvec, b + tile_vecs);
    for (u64 i = b + threadIdx.x; i < e; i += 256 * U) {
      uint4 v[U];
#pragma unroll
      for (int j = 0; j < U; ++j) v[j] = (i + 256ull * j < e) ? ldg_stream(m + i + 256ull * j) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int j = 0; j < U; ++j) acc += v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
    }
  }
  if (acc == 0x12345678u) out[0] = acc;
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// bulk: warp 8 = producer (static contiguous range per CTA), 8 consumer warps xor-reduce the tiles
__global__ void __launch_bounds__(288, 1) k_bulk(const unsigned char *m, u64 nbytes, uint32_t tile_bytes, uint32_t stages, uint32_t *ctr, int dyn, uint32_t *out) {
  extern __shared__ __align__(128) unsigned char smem[];
  u64 *full = (u64 *)(smem + (size_t)stages * tile_bytes), *empty = full + stages;
  uint32_t *s_tile = (uint32_t *)(empty + stages);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (uint32_t s = 0; s < stages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[s])), "r"(8));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const u64 n_tiles = (nbytes + tile_bytes - 1) / tile_bytes;
  const u64 per = (n_tiles + gridDim.x - 1) / gridDim.x;
  const u64 t0 = per * blockIdx.x, t1 = min(n_tiles, t0 + per);
  auto wait = [](u64 *bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
  };
  if (warp == 8) {
    if (lane == 0) {
      u64 pol;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
      uint32_t it = 0;
      for (;; ++it) {
        u64 t;
        if (dyn) t = atomicAdd(ctr, 1u); else t = t0 + it;
        const uint32_t s = it % stages, ph = (it / stages) & 1;
        wait(&empty[s], ph ^ 1);
        const bool done = dyn ? t >= n_tiles : t >= t1;
        s_tile[s] = done ? 0xFFFFFFFFu : (uint32_t)t;
        if (done) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[s])) : "memory"); break; }
        const u64 off = t * tile_bytes;
        const uint32_t bytes = (uint32_t)min((u64)tile_bytes, nbytes - off);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem + (size_t)s * tile_bytes)), "l"(m + off), "r"(bytes), "r"(smem_u32(&full[s])), "l"(pol) : "memory");
      }
    }
    return;
  }
  uint32_t acc = 0;
  for (uint32_t it = 0;; ++it) {
    const uint32_t s = it % stages, ph = (it / stages) & 1;
    wait(&full[s], ph);
    if (s_tile[s] == 0xFFFFFFFFu) break;
    const uint4 *tile = (const uint4 *)(smem + (size_t)s * tile_bytes);
    for (uint32_t v = tid; v < tile_bytes / 16; v += 256) { uint4 x = tile[v]; acc += x.x ^ x.y ^ x.z ^ x.w; }
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
  }
  if (acc == 0x12345678u) out[0] = acc;
}
template <typename F>
float timeit(F f, int iters = 20) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) f();
  cudaEventRecord(a);
  for (int i = 0; i < iters; ++i) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
  return ms / iters;
}
int main() {
  const u64 bytes = 1536000000ull, nvec = bytes / 16;
  unsigned char *d; uint32_t *out, *ctr;
  cudaMalloc(&d, bytes); cudaMalloc(&out, 4); cudaMalloc(&ctr, 4);
  cudaMemset(d, 1, bytes);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  auto rep = [&](const char *name, int a, int b, float ms) { printf("%-10s cfg=(%d,%d)  %.1f us  %.0f GB/s\n", name, a, b, ms * 1e3, bytes / ms / 1e6); fflush(stdout); };
  for (int cps : {1, 2, 3, 4, 6, 8}) {
    rep("static U4", cps, 4, timeit([&] { k_static<4><<<sms * cps, 256>>>((const uint4 *)d, nvec, out); }));
    rep("static U8", cps, 8, timeit([&] { k_static<8><<<sms * cps, 256>>>((const uint4 *)d, nvec, out); }));
    rep("static U12", cps, 12, timeit([&] { k_static<12><<<sms * cps, 256>>>((const uint4 *)d, nvec, out); }));
    rep("stride U4", cps, 4, timeit([&] { k_stride<4><<<sms * cps, 256>>>((const uint4 *)d, nvec, out); }));
    rep("stride U8", cps, 8, timeit([&] { k_stride<8><<<sms * cps, 256>>>((const uint4 *)d, nvec, out); }));
    for (uint32_t tv : {1536u, 3072u, 6144u, 12288u}) {
      rep("dynamic U8", cps, (int)tv, timeit([&] { cudaMemsetAsync(ctr, 0, 4); k_dynamic<8><<<sms * cps, 256>>>((const uint4 *)d, nvec, ctr, tv, out); }));
      rep("dynamic U12", cps, (int)tv, timeit([&] { cudaMemsetAsync(ctr, 0, 4); k_dynamic<12><<<sms * cps, 256>>>((const uint4 *)d, nvec, ctr, tv, out); }));
    }
  }
  cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  for (uint32_t tb : {16384u, 32768u, 49152u}) for (uint32_t st : {3u, 4u, 6u}) for (int dyn : {0, 1}) {
    if ((size_t)tb * st > 200 * 1024) continue;
    size_t sm = (size_t)tb * st + 16 * st + 4 * st + 64;
    rep(dyn ? "bulk dyn" : "bulk stat", (int)tb, (int)st, timeit([&] { cudaMemsetAsync(ctr, 0, 4); k_bulk<<<sms, 288, sm>>>(d, bytes, tb, st, ctr, dyn, out); }));
  }
  // copy reference (same method as MEASURED_PEAKS: read+write bytes)
  unsigned char *d2; cudaMalloc(&d2, bytes);
  rep("memcpy d2d", 0, 0, timeit([&] { cudaMemcpyAsync(d2, d, bytes, cudaMemcpyDeviceToDevice); }) / 2);
  return 0;
}
