// internal.h — private declarations shared by the translation units of libopenintel_gpu.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <condition_variable>
#include <mutex>
#include <string>

#include "../../include/openintel_gpu.h"

typedef unsigned long long u64;

// ---------------------------------------------------------------------------------------------
// per-launch parameter blocks
// ---------------------------------------------------------------------------------------------
struct OiScanParams {
  const uint4 *mat;    // embedding rows, row-major, nv 16-byte vectors per row
  const float *q;      // one f32 query [dim] (device)
  uint32_t n_rows;
  uint32_t nv;         // 16-byte vectors per row
  uint32_t dim;
  uint32_t doc_base;   // added to the local row index when forming keys
  uint32_t k;
  uint32_t rows_per_cta;
  u64 *cand;           // [grid][k] per-CTA sorted candidates
  u64 *gthr;           // grid-wide lower bound of the k-th best key (reset to 0 by the last CTA)
  uint32_t *ticket;    // last-CTA election counter (reset to 0 by the last CTA)
  uint32_t *tile_ctr;  // dynamic tile counter of the bulk variant (reset to 0 by the last CTA)
  u64 *out_keys;       // [k] final sorted keys (0 = empty)
  // batch form (bulk variant: one persistent launch walks nq queries; pointers above = query 0)
  uint32_t nq;
  uint32_t cand_stride;  // keys between two queries' candidate areas
};

struct OiCosineWorkspace {
  u64 *cand = nullptr;       // [max_batch][max_grid][k_stride]
  u64 *gthr = nullptr;       // [max_batch]
  uint32_t *ticket = nullptr;  // [max_batch]
  uint32_t *tile_ctr = nullptr;  // [max_batch]
  uint32_t max_grid = 0;
  uint32_t k_stride = 0;     // = max_k of the index
};

// cosine_scan.cu ------------------------------------------------------------------------------
// Scans the shard for each of nq queries, producing nq sorted key lists d_out_keys[nq][k].
// variant: 0 = direct 128-bit loads, 1 = bulk-async-copy (TMA 1-D) pipeline.
cudaError_t oi_launch_cosine_scan(const void *d_mat, uint32_t dtype, uint64_t n_rows, uint32_t dim,
                                  uint32_t doc_base, const float *d_queries, uint32_t nq, uint32_t k,
                                  const OiCosineWorkspace &ws, u64 *d_out_keys, int variant, int num_sms,
                                  cudaStream_t stream, uint64_t *launches, int multi_query = 0);
uint32_t oi_cosine_scan_max_grid(int num_sms);
void oi_cosine_scan_tuning(int tile_rows, int stages);  // experiments: 0 = default shape

// keys -> (ids, scores) / (ids, rrf ...) unpack; merge of gathered shard lists -----------------
cudaError_t oi_launch_unpack_keys(const u64 *d_keys, uint32_t n, uint32_t *d_ids, float *d_scores,
                                  cudaStream_t stream, uint64_t *launches);
// d_gathered: [world][nq][k] keys -> d_out [nq][k] best k by key.  rank_stride = distance in keys between two
// ranks' blocks (0 = nq * k: the blocks are dense); d_known (optional, [nq]): per query a key that at least k keys
// of its lists are known to reach (0 = none) -- lets the merge copy only the list prefixes that can matter
cudaError_t oi_launch_merge_shards(const u64 *d_gathered, uint32_t world, uint32_t nq, uint32_t k,
                                   u64 *d_out, cudaStream_t stream, uint64_t *launches, size_t rank_stride = 0,
                                   const u64 *d_known = nullptr);

// rrf.cu --------------------------------------------------------------------------------------
cudaError_t oi_launch_rrf(const u64 *d_cos_keys, const u64 *d_bm25_keys, uint32_t nq, uint32_t k, uint32_t rrf_k,
                          uint32_t *d_ids, float *d_rrf, uint32_t *d_rc, uint32_t *d_rb, cudaStream_t stream,
                          uint64_t *launches);

// synth.cu ------------------------------------------------------------------------------------
cudaError_t oi_launch_synth_embeddings(void *d_mat, uint32_t dtype, uint64_t n_rows, uint32_t dim,
                                       uint64_t seed, uint64_t stream_id, uint64_t first_row,
                                       cudaStream_t stream, uint64_t *launches);
