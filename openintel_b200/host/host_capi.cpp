// host_capi.cpp — extern "C" shim over the header-only host layer so that the CPU test-suite can
// drive the tokenizer and the index builder (libopenintel_host.so; no CUDA in here).
#include "openintel_host.hpp"
#include "openintel_store.hpp"

using openintel::IndexBuilder;

extern "C" {

void *oih_builder_create() { return new IndexBuilder(); }
void oih_builder_destroy(void *b) { delete static_cast<IndexBuilder *>(b); }

// posts i = texts[offsets[i] .. offsets[i+1])
void oih_builder_add(void *b, const uint8_t *texts, const uint64_t *offsets, uint64_t n_posts) {
  IndexBuilder *ix = static_cast<IndexBuilder *>(b);
  for (uint64_t i = 0; i < n_posts; ++i)
    ix->add(std::string_view(reinterpret_cast<const char *>(texts) + offsets[i], offsets[i + 1] - offsets[i]));
}
void oih_builder_finish(void *b, uint32_t *n_docs, uint32_t *n_terms, uint64_t *n_postings) {
  IndexBuilder *ix = static_cast<IndexBuilder *>(b);
  ix->finish();
  *n_docs = ix->n_docs();
  *n_terms = ix->n_terms();
  *n_postings = ix->doc_ids().size();
}
void oih_builder_export(void *b, uint64_t *term_offsets, uint32_t *doc_ids, uint32_t *tfs, uint32_t *doc_len) {
  IndexBuilder *ix = static_cast<IndexBuilder *>(b);
  ix->finish();
  std::memcpy(term_offsets, ix->term_offsets().data(), ix->term_offsets().size() * sizeof(uint64_t));
  if (!ix->doc_ids().empty()) {
    std::memcpy(doc_ids, ix->doc_ids().data(), ix->doc_ids().size() * sizeof(uint32_t));
    std::memcpy(tfs, ix->tfs().data(), ix->tfs().size() * sizeof(uint32_t));
  }
  if (ix->n_docs()) std::memcpy(doc_len, ix->doc_len().data(), (size_t)ix->n_docs() * sizeof(uint32_t));
}
uint32_t oih_builder_term_id(void *b, const uint8_t *tok, uint32_t len) {
  return static_cast<IndexBuilder *>(b)->term_id(std::string_view(reinterpret_cast<const char *>(tok), len));
}
// returns the number of term ids the query text maps to (only the first `cap` are written)
uint32_t oih_builder_query_terms(void *b, const uint8_t *text, uint32_t len, uint32_t *out, uint32_t cap) {
  const std::vector<uint32_t> t = static_cast<IndexBuilder *>(b)->query_terms(std::string_view(reinterpret_cast<const char *>(text), len));
  for (uint32_t i = 0; i < t.size() && i < cap; ++i) out[i] = t[i];
  return (uint32_t)t.size();
}
// vocabulary entry `id` -> bytes; returns its length (only the first `cap` bytes are written)
uint32_t oih_builder_term(void *b, uint32_t id, uint8_t *out, uint32_t cap) {
  const std::string &s = static_cast<IndexBuilder *>(b)->vocabulary().at(id);
  std::memcpy(out, s.data(), std::min<size_t>(cap, s.size()));
  return (uint32_t)s.size();
}
// tokens joined by '\n'; returns the joined length (only the first `cap` bytes are written)
uint64_t oih_tokenize(const uint8_t *text, uint64_t len, uint8_t *out, uint64_t cap) {
  std::string joined;
  openintel::tokenize(std::string_view(reinterpret_cast<const char *>(text), len), [&](std::string_view t) {
    if (!joined.empty()) joined.push_back('\n');
    joined.append(t);
  });
  std::memcpy(out, joined.data(), std::min<uint64_t>(cap, joined.size()));
  return joined.size();
}

// ---- SQLite post store (openintel_store.hpp); errors come back as -1 + oih_last_error() ---------
static thread_local std::string g_err;
const char *oih_last_error() { return g_err.c_str(); }
void *oih_store_open(const char *path) {
  try {
    return new openintel::SqlitePostStore(path);
  } catch (const std::exception &e) {
    g_err = e.what();
    return nullptr;
  }
}
void oih_store_close(void *s) { delete static_cast<openintel::SqlitePostStore *>(s); }
void oih_store_info(void *s, uint64_t *n_posts, uint32_t *dim) {
  const auto *st = static_cast<openintel::SqlitePostStore *>(s);
  *n_posts = st->n_posts();
  *dim = st->dim();
}
// adds every post of the store to the builder, doc_id order; returns the number of posts or -1
int64_t oih_store_lift_posts(void *s, void *b) {
  try {
    IndexBuilder *ix = static_cast<IndexBuilder *>(b);
    int64_t n = 0;
    static_cast<openintel::SqlitePostStore *>(s)->for_each_post([&](const openintel::SocialPost &p) { ix->add(p); ++n; });
    return n;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}
// out: [n_posts][dim] f32, rows L2-normalised; returns 0 or -1
int oih_store_embeddings(void *s, float *out) {
  try {
    const std::vector<float> rows = static_cast<openintel::SqlitePostStore *>(s)->embeddings();
    std::memcpy(out, rows.data(), rows.size() * sizeof(float));
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}
// post id of document `doc` as added through oih_store_lift_posts; returns its length
uint32_t oih_builder_post_id(void *b, uint32_t doc, uint8_t *out, uint32_t cap) {
  const std::string &s = static_cast<IndexBuilder *>(b)->post_ids().at(doc);
  std::memcpy(out, s.data(), std::min<size_t>(cap, s.size()));
  return (uint32_t)s.size();
}

// fusion signals (openintel_host.hpp: crowding / alignment / confidence) with the default EngineConfig
double oih_crowding(uint64_t total, double speculation_index, int has_market, int has_rvol, double rvol, int has_iv, double iv_rank) {
  using namespace openintel;
  SocialSummary s;
  s.total = total;
  s.speculation_index = speculation_index;
  MarketSummary m;
  m.has_rvol = has_rvol != 0; m.rvol = rvol; m.has_iv_rank = has_iv != 0; m.iv_rank = iv_rank;
  return crowding(s, has_market ? &m : nullptr);
}
int oih_alignment(uint64_t total, double net_sentiment, int has_market, double pct_change) {
  using namespace openintel;
  SocialSummary s;
  s.total = total;
  s.net_sentiment = net_sentiment;
  MarketSummary m;
  m.pct_change = pct_change;
  return (int)alignment(s, has_market ? &m : nullptr);
}
int oih_confidence(uint64_t n, uint64_t low, uint64_t high) { return (int)openintel::confidence(n, low, high); }

}  // extern "C"
