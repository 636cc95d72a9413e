/* oracle/oracle.h — CPU oracle for the hybrid retrieval path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker.  The product (libopenintel_gpu.so) never links,
 * loads or calls it.
 *
 * PARITY STATUS
 *   - BM25 / cosine / top-k / RRF: "parity unpinned".  The reference (/root/reference) has no
 *     such code, so these functions restate docs/SPEC.md, not reference files.
 *   - tokenizer / lexicon scorer / social summary / crowding / alignment: PINNED to the
 *     reference's own goldens (tests/test_oracle_lexicon.py), each function cites file:line.
 */
#ifndef OIO_ORACLE_H
#define OIO_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OIO_NO_DOC 0xFFFFFFFFu

/* ---- SPEC §9: counter hash and synthetic data -------------------------------------------- */
uint64_t oio_hash64(uint64_t seed, uint64_t stream, uint64_t row, uint64_t col);
/* rows [first_row, first_row+n_rows) of stream `stream`, L2-normalised, f32 row-major */
void oio_synth_rows_f32(uint64_t seed, uint64_t stream, uint64_t first_row, uint64_t n_rows,
                        uint32_t dim, float *out);
/* same rows, RNE-rounded to bf16 (raw u16 bit patterns) */
void oio_synth_rows_bf16(uint64_t seed, uint64_t stream, uint64_t first_row, uint64_t n_rows,
                         uint32_t dim, uint16_t *out);
/* planted queries: v = 2*v_doc[target] + v_noise, target = h(seed,4,j,0) % n_docs */
void oio_synth_planted_queries_f32(uint64_t seed, uint64_t first_q, uint64_t nq, uint32_t dim,
                                   uint64_t n_docs, float *out, uint32_t *out_target);
uint16_t oio_f32_to_bf16(float x);
float oio_bf16_to_f32(uint16_t x);

void oio_synth_doc_lens(uint64_t seed, uint64_t first_doc, uint64_t n_docs, uint32_t *out);
void oio_zipf_cdf(uint32_t vocab, double *cdf);
/* tokens of docs [first_doc, first_doc+n_docs): concatenated, doc d starts at tok_off[d] */
void oio_synth_tokens(uint64_t seed, uint64_t first_doc, uint64_t n_docs, const uint32_t *doc_len,
                      const uint64_t *tok_off, const double *cdf, uint32_t vocab, uint32_t *tokens);
/* query term ids: nq x terms_per_query distinct terms; stream 5 = Zipf, 6 = uniform */
void oio_synth_query_terms(uint64_t seed, uint64_t stream, uint64_t first_q, uint64_t nq,
                           uint32_t terms_per_query, const double *cdf, uint32_t vocab,
                           uint32_t *out);

/* ---- CSR inverted index build (SPEC §3 layout) ------------------------------------------- */
/* pass 1: df per term (term_counts[vocab], zeroed by callee); returns total postings */
uint64_t oio_csr_count(uint64_t n_docs, const uint64_t *tok_off, const uint32_t *tokens,
                       uint32_t vocab, uint64_t *term_counts);
/* pass 2: term_offsets[vocab+1] = exclusive scan of counts; fills doc_ids/tfs (doc ascending) */
void oio_csr_fill(uint64_t n_docs, const uint64_t *tok_off, const uint32_t *tokens, uint32_t vocab,
                  const uint64_t *term_counts, uint64_t *term_offsets, uint32_t *doc_ids,
                  uint32_t *tfs);

/* ---- SPEC §3: BM25 ------------------------------------------------------------------------- */
void oio_bm25_idf(uint64_t n_docs_global, const uint32_t *df, uint32_t n_terms, float *idf);
float oio_bm25_avgdl(const uint32_t *doc_len, uint64_t n_docs);
void oio_bm25_weights(const uint64_t *term_offsets, const uint32_t *doc_ids, const uint32_t *tfs,
                      const uint32_t *doc_len, const float *idf, uint32_t n_terms, float k1,
                      float b, float avgdl, float *w);
/* dense scores of one query (scores[n_docs], overwritten) */
void oio_bm25_score_dense(const uint64_t *term_offsets, const uint32_t *doc_ids, const float *w,
                          uint32_t n_terms, const uint32_t *q_terms, uint32_t n_q_terms,
                          uint64_t n_docs, float *scores);

/* ---- SPEC §1/§2: keys, cosine, top-k ------------------------------------------------------- */
uint64_t oio_key(float score, uint32_t doc_id);
/* double-accumulated dot of f32 rows with an f32 query */
void oio_cosine_scores_f32(const float *rows, uint64_t n, uint32_t dim, const float *q,
                           double *scores);
/* rows bf16; the query is rounded to bf16 first; double accumulation */
void oio_cosine_scores_bf16(const uint16_t *rows, uint64_t n, uint32_t dim, const float *q,
                            double *scores);
/* exact top-k of double scores by (score desc, id asc); ids are doc_base + index */
uint32_t oio_topk_f64(const double *scores, uint64_t n, uint32_t k, uint32_t doc_base,
                      uint32_t *out_ids, double *out_scores);
/* exact top-k of f32 scores by SPEC §1 key; only_positive drops score <= 0; pads to k */
uint32_t oio_topk_f32(const float *scores, uint64_t n, uint32_t k, int only_positive,
                      uint32_t doc_base, uint32_t *out_ids, float *out_scores);
/* f32-accumulated multi-threaded scan + top-k: the CPU *baseline* (bench.py), not the checker */
uint32_t oio_cosine_topk_f32_fast(const float *rows, uint64_t n, uint32_t dim, const float *q,
                                  uint32_t k, int n_threads, uint32_t *out_ids, float *out_scores);
int oio_max_threads(void);

/* ---- SPEC §4: RRF --------------------------------------------------------------------------- */
uint32_t oio_rrf(const uint32_t *ids_cos, uint32_t n_cos, const uint32_t *ids_bm25, uint32_t n_bm25,
                 uint32_t k, uint32_t rrf_k, uint32_t *out_ids, float *out_rrf,
                 uint32_t *out_rank_cos, uint32_t *out_rank_bm25);

/* ---- SPEC §6/§8: tokenizer and lexicon (pinned to the reference) ------------------------- */
/* normalised bytes go to out_norm (capacity >= len); token i = out_norm[tok_start[i] .. +tok_len[i]) */
uint32_t oio_tokenize(const uint8_t *text, size_t len, uint8_t *out_norm, uint32_t *tok_start,
                      uint32_t *tok_len, uint32_t max_tokens);
void oio_lexicon_score(const uint8_t *text, size_t len, double *polarity, int *speculative,
                       uint32_t *bull_hits, uint32_t *bear_hits);
typedef struct {
  uint64_t total, bullish, bearish, neutral;
  double net_sentiment, speculation_index, bull_bear_ratio; /* ratio = -1 when bearish == 0 */
} oio_social_summary_t;
void oio_social_summary(const double *polarity, const int *speculative, uint64_t n,
                        double bull_bear_threshold, oio_social_summary_t *out);
/* has_* flags say whether the optional market component is present */
double oio_crowding(uint64_t total_mentions, double speculation_index, int has_rvol, double rvol,
                    int has_iv, double iv_rank, double w_spec, double w_rvol, double w_iv,
                    double rvol_cap);
/* 0 ConfirmingBullish, 1 ConfirmingBearish, 2 Diverging, 3 Quiet */
int oio_alignment(int has_market, uint64_t total_mentions, double net_sentiment, double pct_change,
                  uint64_t min_sample, double net_thr, double price_thr);

#ifdef __cplusplus
}
#endif
#endif
