/* oracle/oracle_fast.c — the CPU *baseline* leg (bench.py cpu_baseline / --impl reference).
 * TEST/BENCH INFRASTRUCTURE ONLY, never loaded by the product.
 *
 * "self-written CPU oracle (no reference implementation exists)" — SURVEY.md §0.  Same
 * semantics as oracle.c (SPEC §1/§2) but f32-accumulated, vectorisable and OpenMP-parallel so
 * that the CPU number beside the GPU number is a fair one.  Compiled with -O3 -fopenmp and FMA
 * contraction allowed; it is timed, not used as the checker.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int oio_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

typedef float v8f __attribute__((vector_size(32), aligned(4)));

/* 4 independent 8-lane accumulators (AVX2 FMA): the scan should be limited by memory, not by
 * the dependency chain of one accumulator */
static inline float dot_f32(const float *restrict a, const float *restrict b, uint32_t dim) {
  v8f s0 = {0}, s1 = {0}, s2 = {0}, s3 = {0};
  uint32_t c = 0;
  for (; c + 32 <= dim; c += 32) {
    s0 += *(const v8f *)(a + c) * *(const v8f *)(b + c);
    s1 += *(const v8f *)(a + c + 8) * *(const v8f *)(b + c + 8);
    s2 += *(const v8f *)(a + c + 16) * *(const v8f *)(b + c + 16);
    s3 += *(const v8f *)(a + c + 24) * *(const v8f *)(b + c + 24);
  }
  for (; c + 8 <= dim; c += 8) s0 += *(const v8f *)(a + c) * *(const v8f *)(b + c);
  v8f t = (s0 + s1) + (s2 + s3);
  float s = ((t[0] + t[4]) + (t[1] + t[5])) + ((t[2] + t[6]) + (t[3] + t[7]));
  for (; c < dim; ++c) s += a[c] * b[c];
  return s;
}

static void heap_push(uint64_t *heap, uint32_t *m, uint32_t k, uint64_t e) {
  if (*m < k) {
    uint32_t c = (*m)++;
    heap[c] = e;
    while (c > 0) {
      uint32_t p = (c - 1) / 2;
      if (heap[p] > heap[c]) { uint64_t t = heap[p]; heap[p] = heap[c]; heap[c] = t; c = p; } else break;
    }
  } else if (e > heap[0]) {
    heap[0] = e;
    uint32_t c = 0;
    for (;;) {
      uint32_t l = 2 * c + 1, r = l + 1, w = c;
      if (l < *m && heap[l] < heap[w]) w = l;
      if (r < *m && heap[r] < heap[w]) w = r;
      if (w == c) break;
      uint64_t t = heap[w]; heap[w] = heap[c]; heap[c] = t; c = w;
    }
  }
}

static int key_cmp_desc(const void *pa, const void *pb) {
  uint64_t a = *(const uint64_t *)pa, b = *(const uint64_t *)pb;
  return (a < b) - (a > b);
}

uint32_t oio_cosine_topk_f32_fast(const float *rows, uint64_t n, uint32_t dim, const float *q,
                                  uint32_t k, int n_threads, uint32_t *out_ids, float *out_scores) {
  if (n_threads < 1) n_threads = 1;
  uint64_t *heaps = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)k * n_threads);
  uint32_t *counts = (uint32_t *)calloc(n_threads, sizeof(uint32_t));
#pragma omp parallel num_threads(n_threads)
  {
#ifdef _OPENMP
    int t = omp_get_thread_num(), nt = omp_get_num_threads();
#else
    int t = 0, nt = 1;
#endif
    uint64_t lo = n * (uint64_t)t / nt, hi = n * (uint64_t)(t + 1) / nt;
    uint64_t *heap = heaps + (size_t)k * t;
    uint32_t m = 0;
    for (uint64_t r = lo; r < hi; ++r) heap_push(heap, &m, k, oio_key(dot_f32(rows + r * dim, q, dim), (uint32_t)r));
    counts[t] = m;
  }
  uint64_t *all = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)k * n_threads);
  uint32_t m = 0;
  for (int t = 0; t < n_threads; ++t) { memcpy(all + m, heaps + (size_t)k * t, sizeof(uint64_t) * counts[t]); m += counts[t]; }
  qsort(all, m, sizeof(uint64_t), key_cmp_desc);
  uint32_t n_out = m < k ? m : k;
  for (uint32_t i = 0; i < n_out; ++i) {
    uint32_t u = (uint32_t)(all[i] >> 32);
    u ^= (u >> 31) ? 0x80000000u : 0xFFFFFFFFu;
    memcpy(&out_scores[i], &u, 4);
    out_ids[i] = 0xFFFFFFFFu - (uint32_t)all[i];
  }
  for (uint32_t i = n_out; i < k; ++i) { out_ids[i] = OIO_NO_DOC; out_scores[i] = 0.0f; }
  free(heaps); free(counts); free(all);
  return n_out;
}

/* ------------------------------------------------------------------------------------------------
 * CPU baseline of the HYBRID step (bench.py cpu_baseline / --impl reference): one batch of nq queries
 * against a slice of the corpus -- cosine over bf16 rows (blocked so that a block of rows is scored
 * for every query while it sits in cache: the CPU analogue of the batched GEMM), BM25 term-at-a-time
 * over the CSR (oracle.c), per-query top-k of both, RRF (oracle.c).  f32 accumulation, FMA allowed,
 * OpenMP over row blocks / queries.  Timed, not a checker.
 * ---------------------------------------------------------------------------------------------- */
typedef unsigned short v8h __attribute__((vector_size(16), aligned(2)));
typedef unsigned int v8u __attribute__((vector_size(32)));

static inline void bf16_row_to_f32(const uint16_t *restrict src, float *restrict dst, uint32_t dim) {
  uint32_t c = 0;
  for (; c + 8 <= dim; c += 8) {
    v8u w = __builtin_convertvector(*(const v8h *)(src + c), v8u) << 16;
    memcpy(dst + c, &w, 32);
  }
  for (; c < dim; ++c) { uint32_t w = (uint32_t)src[c] << 16; memcpy(dst + c, &w, 4); }
}

/* 4 rows x 1 query: the query vector is loaded once for four dot products */
static inline void dot4_f32(const float *restrict r0, const float *restrict r1, const float *restrict r2,
                            const float *restrict r3, const float *restrict q, uint32_t dim, float out[4]) {
  v8f a0 = {0}, a1 = {0}, a2 = {0}, a3 = {0};
  uint32_t c = 0;
  for (; c + 8 <= dim; c += 8) {
    const v8f x = *(const v8f *)(q + c);
    a0 += *(const v8f *)(r0 + c) * x;
    a1 += *(const v8f *)(r1 + c) * x;
    a2 += *(const v8f *)(r2 + c) * x;
    a3 += *(const v8f *)(r3 + c) * x;
  }
  const v8f *acc[4] = {&a0, &a1, &a2, &a3};
  const float *rows[4] = {r0, r1, r2, r3};
  for (int i = 0; i < 4; ++i) {
    const v8f t = *acc[i];
    float s = ((t[0] + t[4]) + (t[1] + t[5])) + ((t[2] + t[6]) + (t[3] + t[7]));
    for (uint32_t cc = c; cc < dim; ++cc) s += rows[i][cc] * q[cc];
    out[i] = s;
  }
}

/* rows: bf16 [n][dim]; q: f32 [nq][dim] (rounded to bf16 here, SPEC §2); CSR + folded weights w of the same n
 * documents; q_terms [nq][tpq].  Outputs [nq][k].  Returns 0 on success. */
int oio_hybrid_batch_fast(const uint16_t *rows, uint64_t n, uint32_t dim, const float *q, uint32_t nq,
                          const uint64_t *term_offsets, const uint32_t *doc_ids, const float *w, uint32_t n_terms,
                          const uint32_t *q_terms, uint32_t tpq, uint32_t k, uint32_t rrf_k, int n_threads,
                          uint32_t *out_ids, float *out_rrf, uint32_t *out_rank_cos, uint32_t *out_rank_bm25) {
  if (n_threads < 1) n_threads = 1;
  if (dim % 8 != 0 || n > 0xFFFFFFF0ull) return 1;
  const uint32_t RB = 16;  /* rows per block: 16 x dim f32 stays in L1/L2 while every query visits it */
  float *qf = (float *)malloc(sizeof(float) * (size_t)nq * dim);
  for (size_t i = 0; i < (size_t)nq * dim; ++i) qf[i] = oio_bf16_to_f32(oio_f32_to_bf16(q[i]));
  uint64_t *heaps = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)k * nq * n_threads);
  uint32_t *cnt = (uint32_t *)calloc((size_t)nq * n_threads, sizeof(uint32_t));
  uint32_t *cos_ids = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)nq * k);
  uint32_t *bm_ids = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)nq * k);
  const uint64_t n_blocks = (n + RB - 1) / RB;
#pragma omp parallel num_threads(n_threads)
  {
#ifdef _OPENMP
    const int t = omp_get_thread_num();
#else
    const int t = 0;
#endif
    float *blk = (float *)aligned_alloc(64, sizeof(float) * (size_t)RB * dim);
    /* ---- cosine leg: row blocks over the threads, per-(thread, query) heaps ---- */
#pragma omp for schedule(dynamic, 64)
    for (uint64_t b = 0; b < n_blocks; ++b) {
      const uint64_t r0 = b * RB;
      const uint32_t nr = (uint32_t)((r0 + RB <= n) ? RB : n - r0);
      for (uint32_t r = 0; r < nr; ++r) bf16_row_to_f32(rows + (r0 + r) * dim, blk + (size_t)r * dim, dim);
      for (uint32_t r = nr; r < RB; ++r) memset(blk + (size_t)r * dim, 0, sizeof(float) * dim);
      for (uint32_t j = 0; j < nq; ++j) {
        uint64_t *heap = heaps + ((size_t)t * nq + j) * k;
        uint32_t *m = cnt + (size_t)t * nq + j;
        const float *qj = qf + (size_t)j * dim;
        for (uint32_t r = 0; r < nr; r += 4) {
          float s[4];
          dot4_f32(blk + (size_t)r * dim, blk + (size_t)(r + 1) * dim, blk + (size_t)(r + 2) * dim, blk + (size_t)(r + 3) * dim, qj, dim, s);
          for (uint32_t u = 0; u < 4 && r + u < nr; ++u) {
            const uint64_t e = oio_key(s[u], (uint32_t)(r0 + r + u));
            if (*m < k || e > heap[0]) heap_push(heap, m, k, e);
          }
        }
      }
    }
    free(blk);
    /* ---- per query: merge the threads' cosine heaps; BM25 leg; fusion ---- */
    float *scores = (float *)malloc(sizeof(float) * (n ? n : 1));
    uint64_t *all = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)k * n_threads);
    float *tmp_sc = (float *)malloc(sizeof(float) * k);
#pragma omp for schedule(dynamic, 1)
    for (uint32_t j = 0; j < nq; ++j) {
      uint32_t m = 0;
      for (int tt = 0; tt < n_threads; ++tt) {
        const uint32_t c = cnt[(size_t)tt * nq + j];
        memcpy(all + m, heaps + ((size_t)tt * nq + j) * k, sizeof(uint64_t) * c);
        m += c;
      }
      qsort(all, m, sizeof(uint64_t), key_cmp_desc);
      for (uint32_t i = 0; i < k; ++i) cos_ids[(size_t)j * k + i] = i < m ? 0xFFFFFFFFu - (uint32_t)all[i] : OIO_NO_DOC;
      oio_bm25_score_dense(term_offsets, doc_ids, w, n_terms, q_terms + (size_t)j * tpq, tpq, n, scores);
      oio_topk_f32(scores, n, k, 1, 0, bm_ids + (size_t)j * k, tmp_sc);
      oio_rrf(cos_ids + (size_t)j * k, k, bm_ids + (size_t)j * k, k, k, rrf_k, out_ids + (size_t)j * k, out_rrf + (size_t)j * k,
              out_rank_cos + (size_t)j * k, out_rank_bm25 + (size_t)j * k);
    }
    free(scores); free(all); free(tmp_sc);
  }
  free(qf); free(heaps); free(cnt); free(cos_ids); free(bm_ids);
  return 0;
}
