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
