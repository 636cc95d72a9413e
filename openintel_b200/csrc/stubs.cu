// stubs.cu — entry points declared in openintel_gpu.h whose implementation has not landed yet.
// They fail loudly (OI_ERR_UNSUPPORTED); nothing falls back to a CPU path.
#include "handle.h"

void oi_bm25_free(oi_index *) {}
void oi_comm_destroy(oi_index *) {}
oi_status oi_comm_gather_merge(oi_index *h, const u64 *, uint32_t, uint32_t, u64 *, cudaStream_t) {
  return h->fail(OI_ERR_UNSUPPORTED, "multi-GPU merge not built yet");
}
#define UNSUP(h, what) ((h) ? (h)->fail(OI_ERR_UNSUPPORTED, what " not built yet") : OI_ERR_INVALID_ARG)
extern "C" {
oi_status oi_index_load_bm25(oi_index *h, const uint64_t *, const uint32_t *, const uint32_t *, const uint32_t *, uint32_t) { return UNSUP(h, "BM25"); }
oi_status oi_index_synth_bm25(oi_index *h, uint64_t, uint32_t, const double *) { return UNSUP(h, "BM25"); }
oi_status oi_index_bm25_local_stats(oi_index *h, uint32_t *, uint64_t *, uint64_t *) { return UNSUP(h, "BM25"); }
oi_status oi_index_bm25_finalize(oi_index *h, const oi_bm25_params *) { return UNSUP(h, "BM25"); }
oi_status oi_index_read_bm25(oi_index *h, uint64_t *, uint32_t *, uint32_t *, uint32_t *, float *) { return UNSUP(h, "BM25"); }
oi_status oi_comm_unique_id(uint8_t *) { return OI_ERR_UNSUPPORTED; }
oi_status oi_index_comm_init(oi_index *h, int32_t, int32_t, const uint8_t *) { return UNSUP(h, "multi-GPU"); }
oi_status oi_search_bm25(oi_index *h, const uint32_t *, const uint32_t *, uint32_t, uint32_t, uint32_t *, float *) { return UNSUP(h, "BM25"); }
oi_status oi_search_hybrid(oi_index *h, const float *, const uint32_t *, const uint32_t *, uint32_t, uint32_t, uint32_t, uint32_t *, float *, uint32_t *, uint32_t *) { return UNSUP(h, "hybrid"); }
oi_status oi_search_bm25_dev(oi_index *h, const uint32_t *, const uint32_t *, uint32_t, uint32_t, uint32_t *, float *, void *) { return UNSUP(h, "BM25"); }
oi_status oi_search_hybrid_dev(oi_index *h, const float *, const uint32_t *, const uint32_t *, uint32_t, uint32_t, uint32_t, uint32_t *, float *, uint32_t *, uint32_t *, void *) { return UNSUP(h, "hybrid"); }
oi_status oi_lexicon_analyze(int32_t, const uint8_t *, const uint64_t *, uint64_t, double *, uint8_t *, uint32_t *, uint32_t *) { return OI_ERR_UNSUPPORTED; }
}
