// rrf.cu — reciprocal rank fusion of the global cosine and BM25 top-k lists (docs/SPEC.md §4).
// One CTA per query: join the two k-lists by doc id in shared memory, rrf = 1/(rrf_k + rank_cos)
// + 1/(rrf_k + rank_bm25) (f32 IEEE divide, one f32 add), order by the SPEC §1 key, cut to k.
// Integer ranks and single f32 operations make the result bit-exact given identical inputs.
// Built with -fmad=false like bm25.cu (nothing here may be contracted).
#include "internal.h"
#include "oi_common.cuh"

namespace {

constexpr int kRrfThreads = 256;
constexpr int kRrfMaxUnion = 2 * OI_MAX_K;  // 2048

// bitonic sort of key[0..n) descending carrying a u32 payload
__device__ __forceinline__ void bitonic_desc_kv(u64 *key, uint32_t *val, uint32_t n, int tid) {
  for (uint32_t k = 2; k <= n; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t i = tid; i < (n >> 1); i += kRrfThreads) {
        const uint32_t l = ((i & ~(j - 1)) << 1) | (i & (j - 1));
        const uint32_t r = l | j;
        const u64 a = key[l], b = key[r];
        const bool desc = (l & k) == 0;
        if ((a < b) == desc) {
          key[l] = b; key[r] = a;
          const uint32_t t = val[l]; val[l] = val[r]; val[r] = t;
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(kRrfThreads) rrf_kernel(const u64 *cos_keys, const u64 *bm_keys, uint32_t k,
                                                          uint32_t rrf_k, uint32_t *out_ids, float *out_rrf,
                                                          uint32_t *out_rc, uint32_t *out_rb) {
  __shared__ u64 s_key[kRrfMaxUnion];
  __shared__ uint32_t s_slot[kRrfMaxUnion];
  __shared__ uint32_t s_id[kRrfMaxUnion];
  __shared__ uint16_t s_rc[kRrfMaxUnion];
  __shared__ uint16_t s_rb[kRrfMaxUnion];
  const int tid = threadIdx.x;
  const uint32_t q = blockIdx.x;
  const u64 *ck = cos_keys + (size_t)q * k;
  const u64 *bk = bm_keys + (size_t)q * k;
  for (uint32_t i = tid; i < k; i += kRrfThreads) {
    const u64 a = ck[i], b = bk[i];
    s_id[i] = a ? oi_key_doc(a) : OI_NO_DOC_U32;
    s_rc[i] = a ? (uint16_t)(i + 1) : 0;
    s_rb[i] = 0;
    s_id[k + i] = b ? oi_key_doc(b) : OI_NO_DOC_U32;
    s_rc[k + i] = 0;
    s_rb[k + i] = b ? (uint16_t)(i + 1) : 0;
  }
  __syncthreads();
  // join: a BM25 entry whose doc is in the cosine list folds into that entry
  for (uint32_t j = tid; j < k; j += kRrfThreads) {
    const uint32_t d = s_id[k + j];
    if (d == OI_NO_DOC_U32) continue;
    for (uint32_t i = 0; i < k; ++i) {
      if (s_id[i] == d) {
        s_rb[i] = (uint16_t)(j + 1);
        s_id[k + j] = OI_NO_DOC_U32;
        break;
      }
    }
  }
  __syncthreads();
  const uint32_t n = oi_next_pow2(2 * k);
  for (uint32_t e = tid; e < n; e += kRrfThreads) {
    u64 key = 0ull;
    if (e < 2 * k && s_id[e] != OI_NO_DOC_U32) {
      const float a = s_rc[e] ? 1.0f / (float)(rrf_k + s_rc[e]) : 0.0f;
      const float b = s_rb[e] ? 1.0f / (float)(rrf_k + s_rb[e]) : 0.0f;
      const float v = a + b;
      key = oi_make_key(v, s_id[e]);
    }
    s_key[e] = key;
    s_slot[e] = e;
  }
  __syncthreads();
  bitonic_desc_kv(s_key, s_slot, n, tid);
  for (uint32_t o = tid; o < k; o += kRrfThreads) {
    const u64 key = s_key[o];
    const size_t at = (size_t)q * k + o;
    if (key) {
      const uint32_t e = s_slot[o];
      out_ids[at] = s_id[e];
      out_rrf[at] = oi_key_score(key);
      out_rc[at] = s_rc[e];
      out_rb[at] = s_rb[e];
    } else {
      out_ids[at] = OI_NO_DOC_U32;
      out_rrf[at] = 0.0f;
      out_rc[at] = 0;
      out_rb[at] = 0;
    }
  }
}

}  // namespace

cudaError_t oi_launch_rrf(const u64 *d_cos_keys, const u64 *d_bm25_keys, uint32_t nq, uint32_t k, uint32_t rrf_k,
                          uint32_t *d_ids, float *d_rrf, uint32_t *d_rc, uint32_t *d_rb, cudaStream_t stream,
                          uint64_t *launches) {
  if (nq == 0) return cudaSuccess;
  rrf_kernel<<<nq, kRrfThreads, 0, stream>>>(d_cos_keys, d_bm25_keys, k, rrf_k, d_ids, d_rrf, d_rc, d_rb);
  if (launches) ++*launches;
  return cudaGetLastError();
}
