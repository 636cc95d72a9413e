/* oracle/oracle_scale.c — OpenMP drivers that run the oracle (oracle.c) at the BASELINE's full sizes.
 * TEST INFRASTRUCTURE ONLY, never loaded by the product.  Parity status as in oracle.h: "parity unpinned"
 * for BM25 / cosine / top-k.
 *
 * Nothing here restates an algorithm: every score is produced by the functions of oracle.c
 * (oio_synth_rows_*, oio_cosine_scores_*, oio_synth_tokens, ...).  What this file adds is the loop
 * structure that makes a 10M-row brute force finish in seconds: rows / documents are regenerated
 * chunk by chunk from the counter hash (SPEC §9), never stored, and chunks are spread over the host
 * threads.  Compiled with -ffp-contract=off like oracle.c.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { double s; uint32_t id; } sd_t;
/* (score desc, id asc): a before b */
static inline int sd_before(sd_t a, sd_t b) { return a.s > b.s || (a.s == b.s && a.id < b.id); }
static int sd_cmp(const void *pa, const void *pb) {
  sd_t a = *(const sd_t *)pa, b = *(const sd_t *)pb;
  return sd_before(a, b) ? -1 : sd_before(b, a) ? 1 : 0;
}
/* heap whose root is the WORST kept element */
static void sd_push(sd_t *heap, uint32_t *m, uint32_t k, sd_t e) {
  if (*m < k) {
    uint32_t c = (*m)++;
    heap[c] = e;
    while (c > 0) {
      uint32_t p = (c - 1) / 2;
      if (sd_before(heap[p], heap[c])) { sd_t t = heap[p]; heap[p] = heap[c]; heap[c] = t; c = p; } else break;
    }
  } else if (sd_before(e, heap[0])) {
    heap[0] = e;
    uint32_t c = 0;
    for (;;) {
      uint32_t l = 2 * c + 1, r = l + 1, w = c;
      if (l < *m && sd_before(heap[w], heap[l])) w = l;
      if (r < *m && sd_before(heap[w], heap[r])) w = r;
      if (w == c) break;
      sd_t t = heap[w]; heap[w] = heap[c]; heap[c] = t; c = w;
    }
  }
}

/* Exact cosine top-k (double accumulation, SPEC §2) of nq queries over synthetic rows
 * [first_row, first_row + n_rows) of stream 0, regenerated in chunks.  bf16 != 0: rows (and the
 * query) are bf16-rounded as on the bf16 path.  out_*: [nq][k], ids are GLOBAL row numbers. */
void oio_scale_cosine_topk(uint64_t seed, uint64_t first_row, uint64_t n_rows, uint32_t dim, int bf16,
                           const float *q, uint32_t nq, uint32_t k, int n_threads,
                           uint32_t *out_ids, double *out_scores) {
  if (n_threads < 1) n_threads = 1;
  const uint64_t CH = 2048;
  const uint64_t n_chunks = (n_rows + CH - 1) / CH;
  sd_t *heaps = (sd_t *)malloc(sizeof(sd_t) * (size_t)k * nq * n_threads);
  uint32_t *cnt = (uint32_t *)calloc((size_t)nq * n_threads, sizeof(uint32_t));
#pragma omp parallel num_threads(n_threads)
  {
#ifdef _OPENMP
    const int t = omp_get_thread_num();
#else
    const int t = 0;
#endif
    void *rows = malloc((size_t)CH * dim * (bf16 ? 2 : 4));
    double *sc = (double *)malloc(sizeof(double) * CH);
#pragma omp for schedule(dynamic, 4)
    for (uint64_t c = 0; c < n_chunks; ++c) {
      const uint64_t r0 = c * CH, n = (r0 + CH <= n_rows) ? CH : n_rows - r0;
      if (bf16) oio_synth_rows_bf16(seed, 0, first_row + r0, n, dim, (uint16_t *)rows);
      else oio_synth_rows_f32(seed, 0, first_row + r0, n, dim, (float *)rows);
      for (uint32_t j = 0; j < nq; ++j) {
        if (bf16) oio_cosine_scores_bf16((const uint16_t *)rows, n, dim, q + (size_t)j * dim, sc);
        else oio_cosine_scores_f32((const float *)rows, n, dim, q + (size_t)j * dim, sc);
        sd_t *heap = heaps + ((size_t)t * nq + j) * k;
        uint32_t *m = cnt + (size_t)t * nq + j;
        for (uint64_t i = 0; i < n; ++i) {
          sd_t e = {sc[i], (uint32_t)(first_row + r0 + i)};
          sd_push(heap, m, k, e);
        }
      }
    }
    free(rows);
    free(sc);
  }
  sd_t *all = (sd_t *)malloc(sizeof(sd_t) * (size_t)k * n_threads);
  for (uint32_t j = 0; j < nq; ++j) {
    uint32_t m = 0;
    for (int t = 0; t < n_threads; ++t) {
      const uint32_t c = cnt[(size_t)t * nq + j];
      memcpy(all + m, heaps + ((size_t)t * nq + j) * k, sizeof(sd_t) * c);
      m += c;
    }
    qsort(all, m, sizeof(sd_t), sd_cmp);
    for (uint32_t i = 0; i < k; ++i) {
      out_ids[(size_t)j * k + i] = i < m ? all[i].id : OIO_NO_DOC;
      out_scores[(size_t)j * k + i] = i < m ? all[i].s : 0.0;
    }
  }
  free(all); free(heaps); free(cnt);
}

/* rows in parallel (same bytes as oio_synth_rows_bf16 / _f32 called once) */
void oio_scale_synth_rows(uint64_t seed, uint64_t stream, uint64_t first_row, uint64_t n_rows, uint32_t dim,
                          int bf16, int n_threads, void *out) {
  if (n_threads < 1) n_threads = 1;
  const uint64_t CH = 1024;
  const uint64_t n_chunks = (n_rows + CH - 1) / CH;
#pragma omp parallel for schedule(dynamic, 4) num_threads(n_threads)
  for (uint64_t c = 0; c < n_chunks; ++c) {
    const uint64_t r0 = c * CH, n = (r0 + CH <= n_rows) ? CH : n_rows - r0;
    if (bf16) oio_synth_rows_bf16(seed, stream, first_row + r0, n, dim, (uint16_t *)out + (size_t)r0 * dim);
    else oio_synth_rows_f32(seed, stream, first_row + r0, n, dim, (float *)out + (size_t)r0 * dim);
  }
}

/* The CSR restricted to a SMALL SET of terms over synthetic documents [first_doc, first_doc + n_docs): the
 * documents' tokens are regenerated (oio_synth_doc_lens / oio_synth_tokens) chunk by chunk and every
 * (document, wanted term) pair is kept with its term frequency.  Each thread owns a contiguous range of
 * documents and collects its pairs in document order, so the final lists (term-major, concatenated in thread
 * order) are ascending in doc id without a sort.  terms[] must be ascending and distinct, at most 64.
 * Two calls: build returns an opaque object, take copies the arrays out and frees it.  Doc ids are LOCAL
 * (0 .. n_docs-1). */
typedef struct {
  uint32_t n_terms, n_threads;
  uint64_t total, sum_doc_len;
  uint32_t **l_doc, **l_ti, **l_tf;  /* per thread */
  uint64_t *l_n;                     /* per thread */
} oio_mini_csr;

void *oio_scale_mini_csr_build(uint64_t seed, uint64_t first_doc, uint64_t n_docs, const double *cdf, uint32_t vocab,
                               const uint32_t *terms, uint32_t n_terms, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (n_terms > 64) return NULL;
  oio_mini_csr *m = (oio_mini_csr *)calloc(1, sizeof(oio_mini_csr));
  m->n_terms = n_terms; m->n_threads = (uint32_t)n_threads;
  m->l_doc = (uint32_t **)calloc(n_threads, sizeof(uint32_t *));
  m->l_ti = (uint32_t **)calloc(n_threads, sizeof(uint32_t *));
  m->l_tf = (uint32_t **)calloc(n_threads, sizeof(uint32_t *));
  m->l_n = (uint64_t *)calloc(n_threads, sizeof(uint64_t));
  /* term -> index in terms[] (+1), 0 = not wanted */
  uint16_t *want = (uint16_t *)calloc(vocab, sizeof(uint16_t));
  for (uint32_t i = 0; i < n_terms; ++i) if (terms[i] < vocab) want[terms[i]] = (uint16_t)(i + 1);
  uint64_t total_len = 0;
  const uint64_t CH = 4096;
#pragma omp parallel num_threads(n_threads) reduction(+ : total_len)
  {
#ifdef _OPENMP
    const int t = omp_get_thread_num(), nt = omp_get_num_threads();
#else
    const int t = 0, nt = 1;
#endif
    const uint64_t lo = n_docs * (uint64_t)t / nt, hi = n_docs * (uint64_t)(t + 1) / nt;
    uint32_t *dl = (uint32_t *)malloc(sizeof(uint32_t) * CH);
    uint64_t *off = (uint64_t *)malloc(sizeof(uint64_t) * (CH + 1));
    uint32_t *tok = (uint32_t *)malloc(sizeof(uint32_t) * CH * 256);  /* doc length <= 256 (SPEC §9) */
    uint32_t *tfc = (uint32_t *)calloc(n_terms ? n_terms : 1, sizeof(uint32_t));
    uint64_t cap = 1 << 16, ln = 0;
    uint32_t *ld = (uint32_t *)malloc(sizeof(uint32_t) * cap), *li = (uint32_t *)malloc(sizeof(uint32_t) * cap),
             *lf = (uint32_t *)malloc(sizeof(uint32_t) * cap);
    for (uint64_t d0 = lo; d0 < hi; d0 += CH) {
      const uint64_t n = (d0 + CH <= hi) ? CH : hi - d0;
      oio_synth_doc_lens(seed, first_doc + d0, n, dl);
      off[0] = 0;
      for (uint64_t d = 0; d < n; ++d) { off[d + 1] = off[d] + dl[d]; total_len += dl[d]; }
      oio_synth_tokens(seed, first_doc + d0, n, dl, off, cdf, vocab, tok);
      for (uint64_t d = 0; d < n; ++d) {
        int any = 0;
        for (uint64_t i = off[d]; i < off[d + 1]; ++i) {
          const uint16_t wi = want[tok[i]];
          if (wi) { tfc[wi - 1]++; any = 1; }
        }
        if (!any) continue;
        if (ln + n_terms > cap) {
          cap *= 2;
          ld = (uint32_t *)realloc(ld, sizeof(uint32_t) * cap);
          li = (uint32_t *)realloc(li, sizeof(uint32_t) * cap);
          lf = (uint32_t *)realloc(lf, sizeof(uint32_t) * cap);
        }
        for (uint32_t i = 0; i < n_terms; ++i)
          if (tfc[i]) { ld[ln] = (uint32_t)(d0 + d); li[ln] = i; lf[ln] = tfc[i]; ++ln; tfc[i] = 0; }
      }
    }
    m->l_doc[t] = ld; m->l_ti[t] = li; m->l_tf[t] = lf; m->l_n[t] = ln;
    free(dl); free(off); free(tok); free(tfc);
  }
  free(want);
  m->sum_doc_len = total_len;
  for (int t = 0; t < n_threads; ++t) m->total += m->l_n[t];
  return m;
}

uint64_t oio_scale_mini_csr_size(const void *obj) { return obj ? ((const oio_mini_csr *)obj)->total : 0; }

/* term_offsets[n_terms + 1], doc_ids / tfs [size]; frees the object */
void oio_scale_mini_csr_take(void *obj, uint64_t *term_offsets, uint32_t *doc_ids, uint32_t *tfs, uint64_t *sum_doc_len) {
  oio_mini_csr *m = (oio_mini_csr *)obj;
  if (!m) return;
  const uint32_t T = m->n_terms, NT = m->n_threads;
  /* cnt[term][thread] -> write cursors, term-major then thread order (= ascending doc ranges) */
  uint64_t *cur = (uint64_t *)calloc((size_t)(T ? T : 1) * NT, sizeof(uint64_t));
  for (uint32_t t = 0; t < NT; ++t)
    for (uint64_t i = 0; i < m->l_n[t]; ++i) cur[(size_t)m->l_ti[t][i] * NT + t]++;
  uint64_t run = 0;
  for (uint32_t i = 0; i < T; ++i) {
    term_offsets[i] = run;
    for (uint32_t t = 0; t < NT; ++t) { const uint64_t c = cur[(size_t)i * NT + t]; cur[(size_t)i * NT + t] = run; run += c; }
  }
  term_offsets[T] = run;
  for (uint32_t t = 0; t < NT; ++t) {
    for (uint64_t i = 0; i < m->l_n[t]; ++i) {
      const uint64_t p = cur[(size_t)m->l_ti[t][i] * NT + t]++;
      doc_ids[p] = m->l_doc[t][i];
      tfs[p] = m->l_tf[t][i];
    }
    free(m->l_doc[t]); free(m->l_ti[t]); free(m->l_tf[t]);
  }
  if (sum_doc_len) *sum_doc_len = m->sum_doc_len;
  free(cur); free(m->l_doc); free(m->l_ti); free(m->l_tf); free(m->l_n); free(m);
}
