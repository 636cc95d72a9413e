// oi_synth.cuh — the SPEC §9 counter hash, host+device.  Must stay bit-identical to
// oracle/oracle.c (tests/test_gpu_cosine.py::test_device_synth_is_bit_identical_to_oracle and
// tests/test_gpu_bm25.py::test_device_synth_csr_is_bit_identical_to_oracle compare the output bit for bit).
#pragma once
#include <stdint.h>

#define OI_GOLD 0x9E3779B97F4A7C15ULL
#define OI_C_ROW 0xD1B54A32D192ED03ULL

__host__ __device__ __forceinline__ uint64_t oi_mix64(uint64_t z) {
  z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL;
  z ^= z >> 27; z *= 0x94D049BB133111EBULL;
  z ^= z >> 31;
  return z;
}
__host__ __device__ __forceinline__ uint64_t oi_stream_base(uint64_t seed, uint64_t stream) {
  return oi_mix64(seed + OI_GOLD * (stream + 1));
}
__host__ __device__ __forceinline__ uint64_t oi_row_key(uint64_t base, uint64_t row) {
  return oi_mix64(base ^ (row * OI_C_ROW));
}
__host__ __device__ __forceinline__ uint64_t oi_cell(uint64_t rk, uint64_t col) {
  return oi_mix64(rk + OI_GOLD * (col + 1));
}
// Irwin-Hall(4) integer component in [-131070, 131070]
__host__ __device__ __forceinline__ int32_t oi_comp_of(uint64_t h) {
  return (int32_t)((h & 0xFFFF) + ((h >> 16) & 0xFFFF) + ((h >> 32) & 0xFFFF) + (h >> 48)) - 131070;
}
