// cosine_gemm.cu — batched cosine scoring on the 5th-generation tensor cores (BASELINE config 4:
// 10M x 768 bf16, 256 queries, top-100), with the top-k selection fused into the epilogue so that
// the nq x n_docs score matrix never exists in HBM.
//
//   S[q, d] = sum_j Q[q, j] * E[d, j]        Q: queries rounded to bf16 (SPEC §2), E: bf16 rows
//
// Mapping onto tcgen05.mma (kind::f16, cta_group::1, M = 128, N = 64, K = 16):
//   * A = 128 queries.  The whole query tile (128 x dim bf16) is written ONCE into tensor memory
//     (tcgen05.st, dim/2 of the 512 columns) and stays there: the A operand never touches shared
//     memory again, which leaves all of shared memory to the document stream.
//   * B = 64 document rows x 64 bf16 (one 128 B-swizzled TMA box, 8 KB) per pipeline stage, read
//     straight from the row-major embedding matrix by cp.async.bulk.tensor (UTMALDG).
//   * D = 128 x 64 f32 accumulators in tensor memory, double buffered (columns 384..511), so the
//     epilogue of document tile t overlaps the MMAs of tile t+1.
//   * a CTA owns one query tile and every n_ranges-th document tile; the n_qt CTAs that share a
//     document tile run side by side, so the second reader hits L2 and HBM is read once.
//   warp 0 = TMA producer, warp 1 = MMA issuer (one thread) + TMEM allocator, warps 2-5 = epilogue.
//
// Epilogue: thread r of the four epilogue warps owns TMEM lane r = query r of the tile.  It pulls
// the 64 scores of its query with tcgen05.ld, takes their maximum and compares it with the
// query's running threshold; only when the maximum passes does it look at individual scores and
// append (score, doc) keys to the (CTA, query) candidate list in global memory.  A list that
// could overflow is compacted to its best k by the warp (bitonic sort in shared memory) and the
// threshold raised.  A first short pass over a sample of the shard gives every query the exact
// k-th best key of the sample as a starting threshold for the main pass, so in the main pass only
// a handful of scores per (CTA, query) ever pass the filter.
//
//   algorithmic bytes per batch = n_docs * dim * 2        flops per batch = 2 * n_docs * dim * nq
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>

#include "handle.h"
#include "oi_common.cuh"
#include "oi_tcgen05.cuh"

struct OiGemm {
  __nv_bfloat16 *d_qb = nullptr;  // [n_qt_max * 128][dim] bf16 query tiles (zero padded)
  u64 *d_cand = nullptr;          // [max_lists][cap]
  uint32_t *d_cnt = nullptr;      // [max_lists]
  u64 *d_keys_a = nullptr;        // [max_batch][max_k] sample-pass result
  u64 *d_thr_a = nullptr;         // [max_batch]
  uint32_t cap = 0;
  size_t max_lists = 0;
  CUtensorMap tmap;
  bool ready = false;
};

namespace {

constexpr int kGemmThreads = 192;
constexpr uint32_t kTileDocs = 64;    // MMA N: documents per tile
constexpr uint32_t kKBlock = 64;      // bf16 elements per box row (128 B)
constexpr uint32_t kStageBytes = kTileDocs * 128;
constexpr uint32_t kStages = 20;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kDCol0 = 384;      // two 64-column accumulators at 384 and 448
constexpr uint32_t kCapMax = 512;     // keys per (CTA, query) candidate list
constexpr uint32_t kMaxDim = 768;     // dim / 2 columns of TMEM hold the query tile

struct GemmParams {
  const __nv_bfloat16 *qb;
  uint32_t n_rows, dim, doc_base, k, nq, n_qt, n_ranges;
  uint32_t tile_begin, tile_end;  // document tiles [begin, end) covered by this launch
  const u64 *thr_in;              // [nq] starting threshold keys (0 = none) or nullptr
  u64 *cand;                      // [grid * 128][cap]
  uint32_t *cand_cnt;             // [grid * 128]
  uint32_t cap;
  float *dump;                    // tests: [nq][n_rows] raw scores, or nullptr
};

__device__ __forceinline__ void warp_bitonic_desc(u64 *buf, uint32_t n, int lane) {
  for (uint32_t k = 2; k <= n; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t i = lane; i < (n >> 1); i += 32) {
        const uint32_t l = ((i & ~(j - 1)) << 1) | (i & (j - 1));
        const uint32_t r = l | j;
        const u64 a = buf[l], b = buf[r];
        const bool desc = (l & k) == 0;
        if ((a < b) == desc) { buf[l] = b; buf[r] = a; }
      }
      __syncwarp();
    }
  }
}

// The warp reduces list[0..n) to its best min(n, k) keys (sorted descending, written back to the
// front of the list).  Returns the k-th best key (0 when fewer than k are held).
__device__ __forceinline__ u64 warp_compact_list(u64 *list, uint32_t n, uint32_t k, u64 *scratch, int lane) {
  const uint32_t np2 = oi_next_pow2(n);
  for (uint32_t i = lane; i < np2; i += 32) scratch[i] = i < n ? __ldcg(list + i) : 0ull;
  __syncwarp();
  warp_bitonic_desc(scratch, np2, lane);
  const uint32_t keep = min(n, k);
  for (uint32_t i = lane; i < keep; i += 32) list[i] = scratch[i];
  const u64 thr = keep == k ? scratch[k - 1] : 0ull;
  __syncwarp();
  return thr;
}

__global__ void __launch_bounds__(kGemmThreads, 1)
    cosine_gemm_kernel(const __grid_constant__ CUtensorMap tmap, const GemmParams p) {
  extern __shared__ unsigned char s_raw[];
  unsigned char *sm = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(s_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char *s_stage = sm;                                                   // kStages x 8 KB, 1024 B aligned
  u64 *s_scratch = reinterpret_cast<u64 *>(sm + kStages * kStageBytes);          // 4 warps x cap keys
  u64 *s_full = s_scratch + 4 * kCapMax;
  u64 *s_empty = s_full + kStages;
  u64 *s_tfull = s_empty + kStages;   // [2] accumulator ready
  u64 *s_tempty = s_tfull + 2;        // [2] accumulator drained
  u64 *s_qready = s_tempty + 2;
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_qready + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t qt = blockIdx.x % p.n_qt, range = blockIdx.x / p.n_qt;
  const uint32_t nkb = p.dim / kKBlock;
  const uint32_t t0 = p.tile_begin + range;

  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < kStages; ++s) { oi_mbar_init(&s_full[s], 1); oi_mbar_init(&s_empty[s], 1); }
    for (uint32_t b = 0; b < 2; ++b) { oi_mbar_init(&s_tfull[b], 1); oi_mbar_init(&s_tempty[b], 4); }
    oi_mbar_init(s_qready, 128);
    oi_mbar_fence_init();
  }
  if (warp == 1) oi_tmem_alloc(s_tmem, kTmemCols);
  oi_tc_fence_before();
  __syncthreads();
  oi_tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp == 0) {
    // ------------------------------- TMA producer ------------------------------------------------
    if (lane == 0) {
      oi_tma_prefetch_desc(&tmap);
      uint32_t it = 0;
      for (uint32_t t = t0; t < p.tile_end; t += p.n_ranges) {
        for (uint32_t kb = 0; kb < nkb; ++kb, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
          oi_mbar_wait(&s_empty[s], ph ^ 1u);
          oi_mbar_expect_tx(&s_full[s], kStageBytes);
          // rows past the end of the shard are zero-filled by the TMA unit
          oi_tma_load_2d(s_stage + (size_t)s * kStageBytes, &tmap, (int32_t)(kb * kKBlock), (int32_t)(t * kTileDocs), &s_full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------------------------
    if (lane == 0) {
      const uint32_t idesc = oi_umma_idesc_bf16(128, kTileDocs);
      oi_mbar_wait(s_qready, 0);
      oi_tc_fence_after();
      uint32_t it = 0, tl = 0;
      for (uint32_t t = t0; t < p.tile_end; t += p.n_ranges, ++tl) {
        const uint32_t b = tl & 1u, bph = (tl >> 1) & 1u;
        oi_mbar_wait(&s_tempty[b], bph ^ 1u);
        oi_tc_fence_after();
        const uint32_t d_addr = tmem + kDCol0 + b * kTileDocs;
        for (uint32_t kb = 0; kb < nkb; ++kb, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
          oi_mbar_wait(&s_full[s], ph);
          oi_tc_fence_after();
          const uint64_t bdesc = oi_umma_smem_desc_sw128(oi_smem_u32(s_stage + (size_t)s * kStageBytes));
#pragma unroll
          for (uint32_t ks = 0; ks < 4; ++ks) {
            // A: 16 bf16 of K = 8 TMEM columns; B: 16 bf16 of K = 32 B inside the swizzle atom
            oi_umma_ts_bf16(d_addr, tmem + (kb * 4 + ks) * 8, bdesc + (uint64_t)(ks * 2), idesc, (kb | ks) != 0 ? 1u : 0u);
          }
          oi_umma_commit(&s_empty[s]);  // the stage is free once these MMAs have read it
        }
        oi_umma_commit(&s_tfull[b]);
      }
    }
  } else {
    // ------------------------------- epilogue: one thread per query -------------------------------
    const uint32_t quarter = (uint32_t)warp & 3u;  // TMEM lanes this warp may touch: 32 * (warp % 4)
    const uint32_t row = quarter * 32 + lane;
    const uint32_t lane_taddr = tmem + ((quarter * 32u) << 16);
    {
      const uint4 *qrow = reinterpret_cast<const uint4 *>(p.qb + ((size_t)qt * 128 + row) * p.dim);
      for (uint32_t c = 0; c < p.dim / 2; c += 32) {
        uint32_t r[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 x = __ldg(qrow + c / 4 + j);
          r[4 * j] = x.x; r[4 * j + 1] = x.y; r[4 * j + 2] = x.z; r[4 * j + 3] = x.w;
        }
        oi_tmem_st32(lane_taddr + c, r);
      }
      oi_tmem_wait_st();
      oi_tc_fence_before();
      oi_mbar_arrive(s_qready);
    }
    const uint32_t q = qt * 128 + row;
    const bool qv = q < p.nq;
    const uint32_t cap = p.cap;
    const size_t list = (size_t)blockIdx.x * 128 + row;
    u64 *buf = p.cand + list * cap;
    u64 *scratch = s_scratch + (size_t)(warp - 2) * kCapMax;
    uint32_t cnt = 0;
    u64 thr_key = (qv && p.thr_in) ? p.thr_in[q] : 0ull;
    float thr_s = thr_key ? oi_key_score(thr_key) : -INFINITY;

    uint32_t tl = 0;
    for (uint32_t t = t0; t < p.tile_end; t += p.n_ranges, ++tl) {
      const uint32_t b = tl & 1u, bph = (tl >> 1) & 1u;
      oi_mbar_wait(&s_tfull[b], bph);
      oi_tc_fence_after();
      uint32_t v[64];
      oi_tmem_ld32(lane_taddr + kDCol0 + b * kTileDocs, v);
      oi_tmem_ld32(lane_taddr + kDCol0 + b * kTileDocs + 32, v + 32);
      oi_tmem_wait_ld();
      oi_tc_fence_before();
      __syncwarp();
      if (lane == 0) oi_mbar_arrive(&s_tempty[b]);  // the accumulator may be overwritten: scores are in registers

      const uint32_t doc0 = t * kTileDocs;
      const uint32_t n_valid = min(kTileDocs, p.n_rows - doc0);
      if (qv) {
        if (p.dump) {
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if ((uint32_t)i < n_valid) p.dump[(size_t)q * p.n_rows + doc0 + i] = __uint_as_float(v[i]);
        }
        float m = -INFINITY;
        if (n_valid == kTileDocs) {
#pragma unroll
          for (int i = 0; i < 64; ++i) m = fmaxf(m, __uint_as_float(v[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 64; ++i) m = fmaxf(m, (uint32_t)i < n_valid ? __uint_as_float(v[i]) : -INFINITY);
        }
        if (m >= thr_s) {
#pragma unroll
          for (int i = 0; i < 64; ++i) {
            const float s = __uint_as_float(v[i]);
            if (s >= thr_s && (uint32_t)i < n_valid) {
              const u64 key = oi_make_key(s, p.doc_base + doc0 + i);
              if (key > thr_key) { buf[cnt] = key; ++cnt; }
            }
          }
        }
      }
      // a list that could overflow during the next tile is reduced to its best k by the whole warp
      uint32_t mask = __ballot_sync(0xFFFFFFFFu, cnt + kTileDocs > cap);
      while (mask) {
        const int l = __ffs(mask) - 1;
        mask &= mask - 1;
        const uint32_t n_l = __shfl_sync(0xFFFFFFFFu, cnt, l);
        u64 *list_l = p.cand + ((size_t)blockIdx.x * 128 + quarter * 32 + l) * cap;
        __syncwarp();
        const u64 thr_l = warp_compact_list(list_l, n_l, p.k, scratch, lane);
        if (lane == l) {
          cnt = min(n_l, p.k);
          if (thr_l > thr_key) { thr_key = thr_l; thr_s = oi_key_score(thr_l); }
        }
      }
    }
    p.cand_cnt[list] = qv ? cnt : 0u;
  }

  oi_tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    oi_tc_fence_after();
    oi_tmem_dealloc(tmem, kTmemCols);
  }
}

// f32 queries -> bf16 (RNE, SPEC §2), padded with zero rows to a multiple of 128
__global__ void gemm_prep_queries_kernel(const float *q, uint32_t nq, uint32_t rows, uint32_t dim, __nv_bfloat16 *qb) {
  const size_t n = (size_t)rows * dim;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / dim;
    qb[i] = r < nq ? __float2bfloat16_rn(q[i]) : __float2bfloat16_rn(0.0f);
  }
}

struct MergeState {
  u64 buf[OI_SEL_CAP];
  u64 thr;
  uint32_t cnt;
};

// One CTA per query: exact top-k over the query's n_ranges candidate lists (+ an optional sorted
// list from an earlier pass).  out[q][k] sorted descending; thr_out[q] = k-th key or 0.
__global__ void __launch_bounds__(256) gemm_merge_kernel(const u64 *cand, const uint32_t *cnts, uint32_t cap, uint32_t n_ranges,
                                                         uint32_t n_qt, uint32_t k, const u64 *prev, u64 *out, u64 *thr_out) {
  __shared__ MergeState S;
  const int tid = threadIdx.x;
  const uint32_t q = blockIdx.x, qt = q / 128, rl = q % 128;
  if (tid == 0) { S.cnt = 0; S.thr = 0ull; }
  __syncthreads();
  const uint32_t n_list = n_ranges * cap;
  const uint32_t total = n_list + (prev ? k : 0u);
  uint32_t base = 0;
  while (base < total) {
    const uint32_t span = min(total - base, (uint32_t)OI_SEL_CAP - S.cnt);
    const u64 thr = S.thr;
    __syncthreads();
    for (uint32_t i = base + tid; i < base + span; i += 256) {
      u64 key;
      if (i < n_list) {
        const uint32_t r = i / cap, j = i % cap;
        const size_t L = ((size_t)r * n_qt + qt) * 128 + rl;
        key = j < cnts[L] ? __ldcg(cand + L * cap + j) : 0ull;
      } else {
        key = prev[(size_t)q * k + (i - n_list)];
      }
      if (key > thr) oi_sel_push(S.buf, &S.cnt, key);
    }
    base += span;
    __syncthreads();
    if (S.cnt > OI_SEL_CAP / 2 || base >= total) oi_sel_compact(S.buf, &S.cnt, &S.thr, k, tid, 256, 0);
  }
  for (uint32_t i = tid; i < k; i += 256) out[(size_t)q * k + i] = i < S.cnt ? S.buf[i] : 0ull;
  if (thr_out && tid == 0) thr_out[q] = S.cnt == k ? S.buf[k - 1] : 0ull;
}

size_t gemm_smem_bytes() {
  return 1024 + (size_t)kStages * kStageBytes + 4 * kCapMax * sizeof(u64) + (2 * kStages + 5) * sizeof(u64) + 16;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

#define GM_CK(call)                                                                              \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess)                                                                       \
      return h->fail(e_ == cudaErrorMemoryAllocation ? OI_ERR_OUT_OF_MEMORY : OI_ERR_CUDA,       \
                     "%s failed: %s", #call, cudaGetErrorString(e_));                            \
  } while (0)

void oi_gemm_free(oi_index *h) {
  OiGemm *g = h->gemm;
  if (!g) return;
  cudaFree(g->d_qb); cudaFree(g->d_cand); cudaFree(g->d_cnt); cudaFree(g->d_keys_a); cudaFree(g->d_thr_a);
  delete g;
  h->gemm = nullptr;
}

bool oi_gemm_eligible(const oi_index *h, uint32_t nq, uint32_t k) {
  const uint32_t cap = h->gemm_cap ? (uint32_t)h->gemm_cap : kCapMax;
  return h->desc.dtype == OI_DTYPE_BF16 && h->desc.dim % kKBlock == 0 && h->desc.dim <= kMaxDim && h->desc.n_docs > 0 &&
         h->desc.n_docs < 0x7FFFFFC0ull && k + kTileDocs <= cap && nq >= 1;
}

// lazily: workspace + the tensor map of the embedding matrix
static oi_status gemm_prepare(oi_index *h) {
  if (h->gemm && h->gemm->ready) return OI_OK;
  if (!h->gemm) h->gemm = new OiGemm();
  OiGemm *g = h->gemm;
  const size_t B = h->desc.max_batch, K = h->desc.max_k, dim = h->desc.dim;
  const size_t n_qt_max = (B + 127) / 128;
  g->cap = kCapMax;
  g->max_lists = (size_t)std::max<size_t>((size_t)h->num_sms, n_qt_max) * 128;
  GM_CK(cudaMalloc(&g->d_qb, n_qt_max * 128 * dim * sizeof(__nv_bfloat16)));
  GM_CK(cudaMalloc(&g->d_cand, g->max_lists * g->cap * sizeof(u64)));
  GM_CK(cudaMalloc(&g->d_cnt, g->max_lists * sizeof(uint32_t)));
  GM_CK(cudaMalloc(&g->d_keys_a, B * K * sizeof(u64)));
  GM_CK(cudaMalloc(&g->d_thr_a, B * sizeof(u64)));

  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  GM_CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) return h->fail(OI_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
  const cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)h->desc.n_docs};
  const cuuint64_t gstride[1] = {(cuuint64_t)dim * sizeof(__nv_bfloat16)};
  const cuuint32_t box[2] = {kKBlock, kTileDocs};
  const cuuint32_t estr[2] = {1, 1};
  CUresult cr = ((EncodeTiledFn)fn)(&g->tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, h->d_emb, gdim, gstride, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return h->fail(OI_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)cr);
  GM_CK(cudaFuncSetAttribute(cosine_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes()));
  g->ready = true;
  return OI_OK;
}

static oi_status gemm_launch(oi_index *h, uint32_t nq, uint32_t n_qt, uint32_t n_ranges, uint32_t k, uint32_t cap,
                             uint32_t tile_begin, uint32_t tile_end, const u64 *thr_in, float *dump, cudaStream_t st) {
  OiGemm *g = h->gemm;
  GemmParams p;
  p.qb = g->d_qb;
  p.n_rows = (uint32_t)h->desc.n_docs; p.dim = h->desc.dim; p.doc_base = (uint32_t)h->desc.doc_base;
  p.k = k; p.nq = nq; p.n_qt = n_qt; p.n_ranges = n_ranges;
  p.tile_begin = tile_begin; p.tile_end = tile_end;
  p.thr_in = thr_in; p.cand = g->d_cand; p.cand_cnt = g->d_cnt; p.cap = cap; p.dump = dump;
  cosine_gemm_kernel<<<n_ranges * n_qt, kGemmThreads, gemm_smem_bytes(), st>>>(g->tmap, p);
  ++h->launches;
  GM_CK(cudaGetLastError());
  return OI_OK;
}

// Shard-local top-k lists d_out_keys[nq][k] (sorted, global doc ids) of nq f32 queries on the device.
oi_status oi_gemm_local_keys(oi_index *h, const float *d_queries, uint32_t nq, uint32_t k, u64 *d_out_keys, float *d_dump,
                             cudaStream_t st) {
  oi_status s = gemm_prepare(h);
  if (s) return s;
  OiGemm *g = h->gemm;
  const uint32_t n_qt = (nq + 127) / 128;
  uint32_t n_ranges = (uint32_t)h->num_sms / n_qt;
  if (n_ranges < 1) n_ranges = 1;
  const uint32_t n_tiles = (uint32_t)((h->desc.n_docs + kTileDocs - 1) / kTileDocs);
  if (n_ranges > n_tiles) n_ranges = n_tiles;
  if ((size_t)n_ranges * n_qt * 128 > g->max_lists) return h->fail(OI_ERR_CUDA, "internal: GEMM list workspace too small");
  const uint32_t cap = h->gemm_cap ? (uint32_t)h->gemm_cap : g->cap;
  const uint32_t rows = n_qt * 128;
  {
    const size_t n = (size_t)rows * h->desc.dim;
    unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)h->num_sms * 8);
    gemm_prep_queries_kernel<<<blocks, 256, 0, st>>>(d_queries, nq, rows, h->desc.dim, g->d_qb);
    ++h->launches;
    GM_CK(cudaGetLastError());
  }
  // sample pass: as many tiles per CTA as fit a list without compaction
  uint32_t spr = (cap - kTileDocs) / kTileDocs;
  if (h->gemm_sample_tiles > 0) spr = std::min<uint32_t>(spr, (uint32_t)h->gemm_sample_tiles);
  uint32_t n_sample = n_ranges * spr;
  if (d_dump || n_sample >= n_tiles) n_sample = n_tiles;
  if ((s = gemm_launch(h, nq, n_qt, n_ranges, k, cap, 0, n_sample, nullptr, d_dump, st))) return s;
  const bool two = n_sample < n_tiles;
  gemm_merge_kernel<<<nq, 256, 0, st>>>(g->d_cand, g->d_cnt, cap, n_ranges, n_qt, k, nullptr, two ? g->d_keys_a : d_out_keys,
                                        two ? g->d_thr_a : nullptr);
  ++h->launches;
  GM_CK(cudaGetLastError());
  if (two) {
    if ((s = gemm_launch(h, nq, n_qt, n_ranges, k, cap, n_sample, n_tiles, g->d_thr_a, nullptr, st))) return s;
    gemm_merge_kernel<<<nq, 256, 0, st>>>(g->d_cand, g->d_cnt, cap, n_ranges, n_qt, k, g->d_keys_a, d_out_keys, nullptr);
    ++h->launches;
    GM_CK(cudaGetLastError());
  }
  return OI_OK;
}
