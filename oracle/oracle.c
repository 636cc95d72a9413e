/* oracle/oracle.c — CPU oracle.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * "parity unpinned" for BM25 / cosine / top-k / RRF: the reference has no such code
 * (SURVEY.md §0); those functions restate docs/SPEC.md section by section.
 * The tokenizer / lexicon / engine-summary functions at the bottom restate reference files and
 * are pinned to the reference's goldens by tests/test_oracle_lexicon.py.
 *
 * Build: see oracle/Makefile.  This file MUST be compiled with -ffp-contract=off: the BM25
 * functions rely on one IEEE f32 operation per source-level operation (SPEC §3).
 */
#include "oracle.h"
#include "../include/oi_synth_tables.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ============================ SPEC §9: hash + synthetic data ============================ */

#define GOLD 0x9E3779B97F4A7C15ULL
#define C_ROW 0xD1B54A32D192ED03ULL

static inline uint64_t mix64(uint64_t z) {
  z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL;
  z ^= z >> 27; z *= 0x94D049BB133111EBULL;
  z ^= z >> 31;
  return z;
}
static inline uint64_t stream_base(uint64_t seed, uint64_t stream) { return mix64(seed + GOLD * (stream + 1)); }
static inline uint64_t row_key(uint64_t base, uint64_t row) { return mix64(base ^ (row * C_ROW)); }
static inline uint64_t cell(uint64_t rk, uint64_t col) { return mix64(rk + GOLD * (col + 1)); }

uint64_t oio_hash64(uint64_t seed, uint64_t stream, uint64_t row, uint64_t col) {
  return cell(row_key(stream_base(seed, stream), row), col);
}

/* Irwin-Hall(4) integer component in [-131070, 131070] */
static inline int32_t comp_of(uint64_t h) {
  return (int32_t)((h & 0xFFFF) + ((h >> 16) & 0xFFFF) + ((h >> 32) & 0xFFFF) + (h >> 48)) - 131070;
}

uint16_t oio_f32_to_bf16(float x) {
  uint32_t u; memcpy(&u, &x, 4);
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((u >> 16) | 0x40); /* NaN stays NaN */
  u += 0x7FFFu + ((u >> 16) & 1u); /* round to nearest even */
  return (uint16_t)(u >> 16);
}
float oio_bf16_to_f32(uint16_t x) {
  uint32_t u = (uint32_t)x << 16; float f; memcpy(&f, &u, 4); return f;
}

static void normalise_row(const int64_t *v, uint32_t dim, float *out) {
  int64_t ss = 0;
  for (uint32_t c = 0; c < dim; ++c) ss += v[c] * v[c];
  double norm = ss > 0 ? sqrt((double)ss) : 1.0;
  for (uint32_t c = 0; c < dim; ++c) out[c] = (float)((double)v[c] / norm);
}

void oio_synth_rows_f32(uint64_t seed, uint64_t stream, uint64_t first_row, uint64_t n_rows,
                        uint32_t dim, float *out) {
  uint64_t base = stream_base(seed, stream);
  int64_t *v = (int64_t *)malloc(sizeof(int64_t) * dim);
  for (uint64_t r = 0; r < n_rows; ++r) {
    uint64_t rk = row_key(base, first_row + r);
    for (uint32_t c = 0; c < dim; ++c) v[c] = comp_of(cell(rk, c));
    normalise_row(v, dim, out + r * dim);
  }
  free(v);
}

void oio_synth_rows_bf16(uint64_t seed, uint64_t stream, uint64_t first_row, uint64_t n_rows,
                         uint32_t dim, uint16_t *out) {
  float *tmp = (float *)malloc(sizeof(float) * dim);
  for (uint64_t r = 0; r < n_rows; ++r) {
    oio_synth_rows_f32(seed, stream, first_row + r, 1, dim, tmp);
    for (uint32_t c = 0; c < dim; ++c) out[r * dim + c] = oio_f32_to_bf16(tmp[c]);
  }
  free(tmp);
}

void oio_synth_planted_queries_f32(uint64_t seed, uint64_t first_q, uint64_t nq, uint32_t dim,
                                   uint64_t n_docs, float *out, uint32_t *out_target) {
  uint64_t dbase = stream_base(seed, 0), qbase = stream_base(seed, 1);
  int64_t *v = (int64_t *)malloc(sizeof(int64_t) * dim);
  for (uint64_t j = 0; j < nq; ++j) {
    uint64_t q = first_q + j;
    uint64_t target = oio_hash64(seed, 4, q, 0) % n_docs;
    uint64_t drk = row_key(dbase, target), qrk = row_key(qbase, q);
    for (uint32_t c = 0; c < dim; ++c)
      v[c] = 2 * (int64_t)comp_of(cell(drk, c)) + (int64_t)comp_of(cell(qrk, c));
    normalise_row(v, dim, out + j * dim);
    if (out_target) out_target[j] = (uint32_t)target;
  }
  free(v);
}

static const uint16_t LEN_TABLE[OI_LEN_TABLE_SIZE] = OI_LEN_TABLE_INIT;

void oio_synth_doc_lens(uint64_t seed, uint64_t first_doc, uint64_t n_docs, uint32_t *out) {
  for (uint64_t d = 0; d < n_docs; ++d)
    out[d] = LEN_TABLE[oio_hash64(seed, 2, first_doc + d, 0) & (OI_LEN_TABLE_SIZE - 1)];
}

void oio_zipf_cdf(uint32_t vocab, double *cdf) {
  double h = 0.0;
  for (uint32_t r = 0; r < vocab; ++r) { h += 1.0 / (double)(r + 1); cdf[r] = h; }
  for (uint32_t r = 0; r < vocab; ++r) cdf[r] = cdf[r] / h;
  cdf[vocab - 1] = 1.0;
}

/* first r with cdf[r] > u */
static inline uint32_t zipf_draw(uint64_t h, const double *cdf, uint32_t vocab) {
  double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
  uint32_t lo = 0, hi = vocab - 1;
  while (lo < hi) {
    uint32_t mid = lo + (hi - lo) / 2;
    if (cdf[mid] > u) hi = mid; else lo = mid + 1;
  }
  return lo;
}

void oio_synth_tokens(uint64_t seed, uint64_t first_doc, uint64_t n_docs, const uint32_t *doc_len,
                      const uint64_t *tok_off, const double *cdf, uint32_t vocab, uint32_t *tokens) {
  uint64_t base = stream_base(seed, 3);
  for (uint64_t d = 0; d < n_docs; ++d) {
    uint64_t rk = row_key(base, first_doc + d);
    for (uint32_t i = 0; i < doc_len[d]; ++i) tokens[tok_off[d] + i] = zipf_draw(cell(rk, i), cdf, vocab);
  }
}

void oio_synth_query_terms(uint64_t seed, uint64_t stream, uint64_t first_q, uint64_t nq,
                           uint32_t terms_per_query, const double *cdf, uint32_t vocab,
                           uint32_t *out) {
  uint64_t base = stream_base(seed, stream);
  for (uint64_t j = 0; j < nq; ++j) {
    uint64_t rk = row_key(base, first_q + j);
    uint32_t *t = out + j * terms_per_query;
    uint64_t ctr = 0;
    for (uint32_t i = 0; i < terms_per_query; ++i) {
      for (;;) {
        uint64_t h = cell(rk, ctr++);
        uint32_t term = (stream == 6) ? (uint32_t)(h % vocab) : zipf_draw(h, cdf, vocab);
        int dup = 0;
        for (uint32_t p = 0; p < i; ++p) dup |= (t[p] == term);
        if (!dup) { t[i] = term; break; }
      }
    }
  }
}

/* ================================ CSR inverted index ==================================== */

static int cmp_u32(const void *a, const void *b) {
  uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
  return (x > y) - (x < y);
}

uint64_t oio_csr_count(uint64_t n_docs, const uint64_t *tok_off, const uint32_t *tokens,
                       uint32_t vocab, uint64_t *term_counts) {
  memset(term_counts, 0, sizeof(uint64_t) * vocab);
  uint64_t total = 0, cap = 0;
  uint32_t *buf = NULL;
  for (uint64_t d = 0; d < n_docs; ++d) {
    uint64_t len = tok_off[d + 1] - tok_off[d];
    if (len > cap) { cap = len * 2; buf = (uint32_t *)realloc(buf, sizeof(uint32_t) * cap); }
    memcpy(buf, tokens + tok_off[d], sizeof(uint32_t) * len);
    qsort(buf, len, sizeof(uint32_t), cmp_u32);
    for (uint64_t i = 0; i < len; ++i)
      if (i == 0 || buf[i] != buf[i - 1]) { term_counts[buf[i]]++; total++; }
  }
  free(buf);
  return total;
}

void oio_csr_fill(uint64_t n_docs, const uint64_t *tok_off, const uint32_t *tokens, uint32_t vocab,
                  const uint64_t *term_counts, uint64_t *term_offsets, uint32_t *doc_ids,
                  uint32_t *tfs) {
  term_offsets[0] = 0;
  for (uint32_t t = 0; t < vocab; ++t) term_offsets[t + 1] = term_offsets[t] + term_counts[t];
  uint64_t *cursor = (uint64_t *)malloc(sizeof(uint64_t) * vocab);
  memcpy(cursor, term_offsets, sizeof(uint64_t) * vocab);
  uint64_t cap = 0;
  uint32_t *buf = NULL;
  for (uint64_t d = 0; d < n_docs; ++d) {
    uint64_t len = tok_off[d + 1] - tok_off[d];
    if (len > cap) { cap = len * 2; buf = (uint32_t *)realloc(buf, sizeof(uint32_t) * cap); }
    memcpy(buf, tokens + tok_off[d], sizeof(uint32_t) * len);
    qsort(buf, len, sizeof(uint32_t), cmp_u32);
    uint64_t i = 0;
    while (i < len) {
      uint64_t j = i;
      while (j < len && buf[j] == buf[i]) ++j;
      uint64_t p = cursor[buf[i]]++;
      doc_ids[p] = (uint32_t)d;
      tfs[p] = (uint32_t)(j - i);
      i = j;
    }
  }
  free(buf);
  free(cursor);
}

/* ==================================== SPEC §3: BM25 ===================================== */

void oio_bm25_idf(uint64_t n_docs_global, const uint32_t *df, uint32_t n_terms, float *idf) {
  for (uint32_t t = 0; t < n_terms; ++t) {
    double d = (double)df[t];
    idf[t] = df[t] == 0 ? 0.0f : (float)log(1.0 + ((double)n_docs_global - d + 0.5) / (d + 0.5));
  }
}

float oio_bm25_avgdl(const uint32_t *doc_len, uint64_t n_docs) {
  uint64_t s = 0;
  for (uint64_t d = 0; d < n_docs; ++d) s += doc_len[d];
  return n_docs ? (float)((double)s / (double)n_docs) : 1.0f;
}

void oio_bm25_weights(const uint64_t *term_offsets, const uint32_t *doc_ids, const uint32_t *tfs,
                      const uint32_t *doc_len, const float *idf, uint32_t n_terms, float k1,
                      float b, float avgdl, float *w) {
  const float one_minus_b = 1.0f - b;
  const float k1p1 = k1 + 1.0f;
  for (uint32_t t = 0; t < n_terms; ++t) {
    for (uint64_t p = term_offsets[t]; p < term_offsets[t + 1]; ++p) {
      float dl = (float)doc_len[doc_ids[p]];
      float ratio = dl / avgdl;
      float bt = b * ratio;
      float u = one_minus_b + bt;
      float norm = k1 * u;
      float tf = (float)tfs[p];
      float num = tf * k1p1;
      float den = tf + norm;
      float q = num / den;
      w[p] = idf[t] * q;
    }
  }
}

void oio_bm25_score_dense(const uint64_t *term_offsets, const uint32_t *doc_ids, const float *w,
                          uint32_t n_terms, const uint32_t *q_terms, uint32_t n_q_terms,
                          uint64_t n_docs, float *scores) {
  memset(scores, 0, sizeof(float) * n_docs);
  uint32_t *t = (uint32_t *)malloc(sizeof(uint32_t) * (n_q_terms ? n_q_terms : 1));
  memcpy(t, q_terms, sizeof(uint32_t) * n_q_terms);
  qsort(t, n_q_terms, sizeof(uint32_t), cmp_u32);
  for (uint32_t i = 0; i < n_q_terms; ++i) {
    if (i > 0 && t[i] == t[i - 1]) continue; /* repeated query terms count once */
    if (t[i] >= n_terms) continue;           /* unknown term */
    for (uint64_t p = term_offsets[t[i]]; p < term_offsets[t[i] + 1]; ++p) {
      float s = scores[doc_ids[p]];
      scores[doc_ids[p]] = s + w[p];
    }
  }
  free(t);
}

/* =============================== SPEC §1/§2: keys, cosine, top-k ========================== */

uint64_t oio_key(float score, uint32_t doc_id) {
  float c = score + 0.0f;
  uint32_t u; memcpy(&u, &c, 4);
  u ^= (u >> 31) ? 0xFFFFFFFFu : 0x80000000u;
  return ((uint64_t)u << 32) | (uint64_t)(0xFFFFFFFFu - doc_id);
}
static inline float key_score(uint64_t key) {
  uint32_t u = (uint32_t)(key >> 32);
  u ^= (u >> 31) ? 0x80000000u : 0xFFFFFFFFu;
  float f; memcpy(&f, &u, 4); return f;
}
static inline uint32_t key_doc(uint64_t key) { return 0xFFFFFFFFu - (uint32_t)key; }

void oio_cosine_scores_f32(const float *rows, uint64_t n, uint32_t dim, const float *q,
                           double *scores) {
  for (uint64_t r = 0; r < n; ++r) {
    const float *row = rows + r * dim;
    double acc = 0.0;
    for (uint32_t c = 0; c < dim; ++c) acc += (double)row[c] * (double)q[c];
    scores[r] = acc;
  }
}

void oio_cosine_scores_bf16(const uint16_t *rows, uint64_t n, uint32_t dim, const float *q,
                            double *scores) {
  double *qd = (double *)malloc(sizeof(double) * dim);
  for (uint32_t c = 0; c < dim; ++c) qd[c] = (double)oio_bf16_to_f32(oio_f32_to_bf16(q[c]));
  for (uint64_t r = 0; r < n; ++r) {
    const uint16_t *row = rows + r * dim;
    double acc = 0.0;
    for (uint32_t c = 0; c < dim; ++c) acc += (double)oio_bf16_to_f32(row[c]) * qd[c];
    scores[r] = acc;
  }
  free(qd);
}

typedef struct { double s; uint32_t id; } sd_t;
/* a ranks before b */
static inline int sd_before(sd_t a, sd_t b) { return a.s > b.s || (a.s == b.s && a.id < b.id); }
static int sd_cmp(const void *pa, const void *pb) {
  sd_t a = *(const sd_t *)pa, b = *(const sd_t *)pb;
  return sd_before(a, b) ? -1 : (sd_before(b, a) ? 1 : 0);
}

uint32_t oio_topk_f64(const double *scores, uint64_t n, uint32_t k, uint32_t doc_base,
                      uint32_t *out_ids, double *out_scores) {
  if (k == 0) return 0;
  sd_t *heap = (sd_t *)malloc(sizeof(sd_t) * k); /* root = worst kept */
  uint32_t m = 0;
  for (uint64_t i = 0; i < n; ++i) {
    sd_t e = {scores[i], doc_base + (uint32_t)i};
    if (m < k) {
      uint32_t c = m++;
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
        if (l < m && sd_before(heap[w], heap[l])) w = l;
        if (r < m && sd_before(heap[w], heap[r])) w = r;
        if (w == c) break;
        sd_t t = heap[w]; heap[w] = heap[c]; heap[c] = t; c = w;
      }
    }
  }
  qsort(heap, m, sizeof(sd_t), sd_cmp);
  for (uint32_t i = 0; i < m; ++i) { out_ids[i] = heap[i].id; out_scores[i] = heap[i].s; }
  for (uint32_t i = m; i < k; ++i) { out_ids[i] = OIO_NO_DOC; out_scores[i] = 0.0; }
  free(heap);
  return m;
}

static int key_cmp_desc(const void *pa, const void *pb) {
  uint64_t a = *(const uint64_t *)pa, b = *(const uint64_t *)pb;
  return (a < b) - (a > b);
}

uint32_t oio_topk_f32(const float *scores, uint64_t n, uint32_t k, int only_positive,
                      uint32_t doc_base, uint32_t *out_ids, float *out_scores) {
  if (k == 0) return 0;
  uint64_t *heap = (uint64_t *)malloc(sizeof(uint64_t) * k); /* min-heap on key */
  uint32_t m = 0;
  for (uint64_t i = 0; i < n; ++i) {
    if (only_positive && !(scores[i] > 0.0f)) continue;
    uint64_t e = oio_key(scores[i], doc_base + (uint32_t)i);
    if (m < k) {
      uint32_t c = m++;
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
        if (l < m && heap[l] < heap[w]) w = l;
        if (r < m && heap[r] < heap[w]) w = r;
        if (w == c) break;
        uint64_t t = heap[w]; heap[w] = heap[c]; heap[c] = t; c = w;
      }
    }
  }
  qsort(heap, m, sizeof(uint64_t), key_cmp_desc);
  for (uint32_t i = 0; i < m; ++i) { out_ids[i] = key_doc(heap[i]); out_scores[i] = key_score(heap[i]); }
  for (uint32_t i = m; i < k; ++i) { out_ids[i] = OIO_NO_DOC; out_scores[i] = 0.0f; }
  free(heap);
  return m;
}

/* ==================================== SPEC §4: RRF ====================================== */

uint32_t oio_rrf(const uint32_t *ids_cos, uint32_t n_cos, const uint32_t *ids_bm25, uint32_t n_bm25,
                 uint32_t k, uint32_t rrf_k, uint32_t *out_ids, float *out_rrf,
                 uint32_t *out_rank_cos, uint32_t *out_rank_bm25) {
  uint32_t cap = n_cos + n_bm25, m = 0;
  uint32_t *ids = (uint32_t *)malloc(sizeof(uint32_t) * (cap ? cap : 1));
  uint32_t *rc = (uint32_t *)calloc(cap ? cap : 1, sizeof(uint32_t));
  uint32_t *rb = (uint32_t *)calloc(cap ? cap : 1, sizeof(uint32_t));
  for (uint32_t i = 0; i < n_cos; ++i) {
    if (ids_cos[i] == OIO_NO_DOC) continue;
    ids[m] = ids_cos[i]; rc[m] = i + 1; ++m;
  }
  for (uint32_t j = 0; j < n_bm25; ++j) {
    if (ids_bm25[j] == OIO_NO_DOC) continue;
    uint32_t at = m;
    for (uint32_t i = 0; i < m; ++i) if (ids[i] == ids_bm25[j]) { at = i; break; }
    if (at == m) { ids[m] = ids_bm25[j]; ++m; }
    rb[at] = j + 1;
  }
  uint64_t *keys = (uint64_t *)malloc(sizeof(uint64_t) * (m ? m : 1));
  uint32_t *slot = (uint32_t *)malloc(sizeof(uint32_t) * (m ? m : 1));
  float *val = (float *)malloc(sizeof(float) * (m ? m : 1));
  for (uint32_t i = 0; i < m; ++i) {
    float a = rc[i] ? 1.0f / (float)(rrf_k + rc[i]) : 0.0f;
    float b = rb[i] ? 1.0f / (float)(rrf_k + rb[i]) : 0.0f;
    val[i] = a + b;
    keys[i] = oio_key(val[i], ids[i]);
  }
  /* selection by key: m <= 2048, quadratic is fine and obviously correct */
  uint32_t n_out = m < k ? m : k;
  char *used = (char *)calloc(m ? m : 1, 1);
  for (uint32_t o = 0; o < n_out; ++o) {
    uint32_t best = 0xFFFFFFFFu;
    for (uint32_t i = 0; i < m; ++i)
      if (!used[i] && (best == 0xFFFFFFFFu || keys[i] > keys[best])) best = i;
    used[best] = 1; slot[o] = best;
  }
  for (uint32_t o = 0; o < n_out; ++o) {
    uint32_t i = slot[o];
    out_ids[o] = ids[i]; out_rrf[o] = val[i]; out_rank_cos[o] = rc[i]; out_rank_bm25[o] = rb[i];
  }
  for (uint32_t o = n_out; o < k; ++o) { out_ids[o] = OIO_NO_DOC; out_rrf[o] = 0.0f; out_rank_cos[o] = 0; out_rank_bm25[o] = 0; }
  free(ids); free(rc); free(rb); free(keys); free(slot); free(val); free(used);
  return n_out;
}

/* ============ SPEC §6: tokenizer — restates src/adapters/analyzer/lexicon.rs:54-58 ======= */
/* `text.to_lowercase()` then split on `!c.is_ascii_alphanumeric()`, empty tokens dropped.
 * Working on UTF-8 bytes: every non-ASCII char is a separator, except the two code points whose
 * Unicode lowercase contains an ASCII letter: U+212A (E2 84 AA) -> "k"; U+0130 (C4 B0) -> "i"
 * followed by U+0307 (a separator). */
uint32_t oio_tokenize(const uint8_t *text, size_t len, uint8_t *out_norm, uint32_t *tok_start,
                      uint32_t *tok_len, uint32_t max_tokens) {
  uint32_t n_tok = 0, w = 0, cur_start = 0, cur_len = 0;
  size_t i = 0;
  while (i <= len) {
    int ch = -1; /* -1 = separator */
    int sep_after = 0;
    if (i < len) {
      uint8_t b = text[i];
      if (b < 0x80) {
        if (b >= 'A' && b <= 'Z') ch = b + 32;
        else if ((b >= 'a' && b <= 'z') || (b >= '0' && b <= '9')) ch = b;
        i += 1;
      } else if (b == 0xE2 && i + 2 < len && text[i + 1] == 0x84 && text[i + 2] == 0xAA) {
        ch = 'k'; i += 3;
      } else if (b == 0xC4 && i + 1 < len && text[i + 1] == 0xB0) {
        ch = 'i'; sep_after = 1; i += 2;
      } else {
        i += 1;
      }
    } else {
      i += 1; /* end of text acts as a final separator */
    }
    if (ch >= 0) {
      if (cur_len == 0) cur_start = w;
      out_norm[w++] = (uint8_t)ch; cur_len++;
    }
    if (ch < 0 || sep_after) {
      if (cur_len > 0) {
        if (n_tok < max_tokens) { tok_start[n_tok] = cur_start; tok_len[n_tok] = cur_len; }
        n_tok++; cur_len = 0;
      }
    }
  }
  return n_tok;
}

/* ===== SPEC §8: lexicon scorer — restates src/adapters/analyzer/lexicon.rs:9-44, 53-73 ===== */
static const char *const BULL[] = {"moon", "calls", "long", "buy", "bullish", "squeeze", "breakout",
                                   "rocket", "pump", "rip", "green", "up", "rally", "bull"};
static const char *const BEAR[] = {"puts", "short", "sell", "bearish", "dump", "crash", "drilling",
                                   "bagholder", "rug", "red", "down", "tank", "bear"};
static const char *const JARGON[] = {"calls", "puts", "0dte", "yolo", "leaps", "theta", "gamma",
                                     "squeeze", "otm", "itm", "strike", "iv", "delta", "vega",
                                     "contracts"};
#define N_OF(a) (sizeof(a) / sizeof((a)[0]))

static int in_list(const uint8_t *tok, uint32_t len, const char *const *list, size_t n) {
  for (size_t i = 0; i < n; ++i)
    if (strlen(list[i]) == len && memcmp(list[i], tok, len) == 0) return 1;
  return 0;
}

void oio_lexicon_score(const uint8_t *text, size_t len, double *polarity, int *speculative,
                       uint32_t *bull_hits, uint32_t *bear_hits) {
  uint8_t *norm = (uint8_t *)malloc(len + 1);
  uint32_t cap = (uint32_t)(len / 1 + 1);
  uint32_t *st = (uint32_t *)malloc(sizeof(uint32_t) * cap), *ln = (uint32_t *)malloc(sizeof(uint32_t) * cap);
  uint32_t n = oio_tokenize(text, len, norm, st, ln, cap);
  uint32_t bull = 0, bear = 0; int spec = 0;
  for (uint32_t i = 0; i < n; ++i) {
    bull += in_list(norm + st[i], ln[i], BULL, N_OF(BULL));
    bear += in_list(norm + st[i], ln[i], BEAR, N_OF(BEAR));
    spec |= in_list(norm + st[i], ln[i], JARGON, N_OF(JARGON));
  }
  double bh = (double)bull, eh = (double)bear;
  double p = (bh + eh == 0.0) ? 0.0 : (bh - eh) / (bh + eh);
  if (p != p) p = 0.0;              /* Polarity::new — src/domain/values/polarity.rs:8-14 */
  if (p > 1.0) p = 1.0;
  if (p < -1.0) p = -1.0;
  *polarity = p; *speculative = spec;
  if (bull_hits) *bull_hits = bull;
  if (bear_hits) *bear_hits = bear;
  free(norm); free(st); free(ln);
}

/* restates SpeculationEngine::social_summary — src/domain/engine/speculation_engine.rs:70-125 */
void oio_social_summary(const double *polarity, const int *speculative, uint64_t n,
                        double bull_bear_threshold, oio_social_summary_t *out) {
  uint64_t bullish = 0, bearish = 0, neutral = 0, spec = 0;
  double sum = 0.0;
  for (uint64_t i = 0; i < n; ++i) {
    double v = polarity[i];
    sum += v;
    if (v > bull_bear_threshold) bullish++;
    else if (v < -bull_bear_threshold) bearish++;
    else neutral++;
    if (speculative[i]) spec++;
  }
  double net = n == 0 ? 0.0 : sum / (double)n;
  double si = n == 0 ? 0.0 : (double)spec / (double)n;
  if (net > 1.0) net = 1.0;
  if (net < -1.0) net = -1.0;
  if (si > 1.0) si = 1.0;
  if (si < 0.0) si = 0.0;
  out->total = n; out->bullish = bullish; out->bearish = bearish; out->neutral = neutral;
  out->net_sentiment = net; out->speculation_index = si;
  out->bull_bear_ratio = bearish == 0 ? -1.0 : (double)bullish / (double)bearish;
}

static inline double clamp01(double x) { return x < 0.0 ? 0.0 : (x > 1.0 ? 1.0 : x); }

/* restates SpeculationEngine::crowding — src/domain/engine/speculation_engine.rs:151-176 */
double oio_crowding(uint64_t total_mentions, double speculation_index, int has_rvol, double rvol,
                    int has_iv, double iv_rank, double w_spec, double w_rvol, double w_iv,
                    double rvol_cap) {
  double weighted = 0.0, wsum = 0.0;
  if (total_mentions > 0) { weighted += w_spec * speculation_index; wsum += w_spec; }
  if (has_rvol) { weighted += w_rvol * clamp01(rvol / rvol_cap); wsum += w_rvol; }
  if (has_iv) { weighted += w_iv * clamp01(iv_rank); wsum += w_iv; }
  return wsum == 0.0 ? 0.0 : clamp01(weighted / wsum);
}

/* restates SpeculationEngine::alignment — src/domain/engine/speculation_engine.rs:178-208 */
int oio_alignment(int has_market, uint64_t total_mentions, double net_sentiment, double pct_change,
                  uint64_t min_sample, double net_thr, double price_thr) {
  if (!has_market) return 3;
  if (total_mentions < min_sample) return 3;
  if (!(fabs(net_sentiment) >= net_thr) || !(fabs(pct_change) >= price_thr)) return 3;
  int sp = net_sentiment > 0.0, pp = pct_change > 0.0;
  if (sp && pp) return 0;
  if (!sp && !pp) return 1;
  return 2;
}
