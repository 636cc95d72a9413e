// synth.cu — on-device generation of the synthetic corpus (docs/SPEC.md §9).
// Counter-hash based, integer until the final normalisation, so every row is bit-identical to
// the CPU oracle's without moving the matrix across PCIe (153.6 GB at config 5).
#include <cuda_bf16.h>

#include "internal.h"
#include "oi_synth.cuh"

namespace {

// One warp per row.  Pass 1: integer sum of squares; pass 2: regenerate, scale, store.
template <typename T>
__global__ void __launch_bounds__(256) synth_rows_kernel(T *mat, uint64_t n_rows, uint32_t dim, uint64_t seed,
                                                         uint64_t stream_id, uint64_t first_row) {
  const uint64_t warp_global = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  const uint64_t base = oi_stream_base(seed, stream_id);
  for (uint64_t r = warp_global; r < n_rows; r += n_warps) {
    const uint64_t rk = oi_row_key(base, first_row + r);
    long long ss = 0;
    for (uint32_t c = lane; c < dim; c += 32) {
      long long v = oi_comp_of(oi_cell(rk, c));
      ss += v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
    const double norm = ss > 0 ? sqrt((double)ss) : 1.0;
    T *row = mat + r * dim;
    for (uint32_t c = lane; c < dim; c += 32) {
      const float x = (float)((double)oi_comp_of(oi_cell(rk, c)) / norm);
      if constexpr (sizeof(T) == 4) {
        row[c] = x;
      } else {
        row[c] = __float2bfloat16_rn(x);
      }
    }
  }
}

}  // namespace

cudaError_t oi_launch_synth_embeddings(void *d_mat, uint32_t dtype, uint64_t n_rows, uint32_t dim,
                                       uint64_t seed, uint64_t stream_id, uint64_t first_row,
                                       cudaStream_t stream, uint64_t *launches) {
  if (n_rows == 0) return cudaSuccess;
  uint64_t blocks = (n_rows + 7) / 8;
  if (blocks > 148ull * 32) blocks = 148ull * 32;
  if (dtype == OI_DTYPE_F32)
    synth_rows_kernel<float><<<(unsigned)blocks, 256, 0, stream>>>((float *)d_mat, n_rows, dim, seed, stream_id, first_row);
  else
    synth_rows_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, stream>>>((__nv_bfloat16 *)d_mat, n_rows, dim, seed, stream_id, first_row);
  if (launches) ++*launches;
  return cudaGetLastError();
}
