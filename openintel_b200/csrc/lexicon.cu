// lexicon.cu — batched GPU PostAnalyzer: the reference's LexiconAnalyzer::score
// (src/adapters/analyzer/lexicon.rs:53-73, word lists :9-44) over a whole batch of posts.
//
// Tokenizer (lexicon.rs:54-58; docs/SPEC.md §6): Unicode-lowercase, split on every char that is
// not ASCII alphanumeric.  On UTF-8 bytes: ASCII letters/digits are token bytes (A-Z lowered),
// every other byte is a separator, except U+212A KELVIN SIGN (E2 84 AA -> 'k', its two
// continuation bytes are transparent) and U+0130 (C4 B0 -> 'i' followed by a separator).
//
// One warp per post; lanes stride over the byte positions.  A lane that sits on the first byte of
// a token gathers its (<= 9 relevant) normalised bytes into a 128-bit word and compares it with
// the packed BULL / BEAR / JARGON tables.  Integer hit counts are exact; polarity is one f64
// divide (Polarity::new clamps to [-1, 1], src/domain/values/polarity.rs:8-14).
// HBM-bound byte scan: algorithmic bytes = the text bytes, read once.
#include <string>

#include "internal.h"

namespace {

struct Word { unsigned long long lo, hi; };

constexpr Word pack_word(const char *s) {
  unsigned long long lo = 0, hi = 0;
  int i = 0;
  for (; s[i] && i < 8; ++i) lo |= (unsigned long long)(unsigned char)s[i] << (8 * i);
  for (; s[i] && i < 16; ++i) hi |= (unsigned long long)(unsigned char)s[i] << (8 * (i - 8));
  return Word{lo, hi};
}

constexpr int N_BULL = 14, N_BEAR = 13, N_JARGON = 15;
__constant__ Word c_bull[N_BULL] = {pack_word("moon"), pack_word("calls"), pack_word("long"), pack_word("buy"),
                                    pack_word("bullish"), pack_word("squeeze"), pack_word("breakout"),
                                    pack_word("rocket"), pack_word("pump"), pack_word("rip"), pack_word("green"),
                                    pack_word("up"), pack_word("rally"), pack_word("bull")};
__constant__ Word c_bear[N_BEAR] = {pack_word("puts"), pack_word("short"), pack_word("sell"), pack_word("bearish"),
                                    pack_word("dump"), pack_word("crash"), pack_word("drilling"),
                                    pack_word("bagholder"), pack_word("rug"), pack_word("red"), pack_word("down"),
                                    pack_word("tank"), pack_word("bear")};
__constant__ Word c_jargon[N_JARGON] = {pack_word("calls"), pack_word("puts"), pack_word("0dte"), pack_word("yolo"),
                                        pack_word("leaps"), pack_word("theta"), pack_word("gamma"),
                                        pack_word("squeeze"), pack_word("otm"), pack_word("itm"), pack_word("strike"),
                                        pack_word("iv"), pack_word("delta"), pack_word("vega"), pack_word("contracts")};

enum { CLS_SEP = 0, CLS_SKIP = -1, CLS_I_THEN_SEP = -2 };  // > 0: a token byte (the lowered char)

// class of byte position i of text[0..len): token char (> 0), separator, transparent, or the
// 'i' of U+0130 (a token char that also ends the token)
__device__ __forceinline__ int classify(const uint8_t *t, long long i, long long len) {
  const uint8_t b = t[i];
  if (b < 0x80) {
    if (b >= 'A' && b <= 'Z') return b + 32;
    if ((b >= 'a' && b <= 'z') || (b >= '0' && b <= '9')) return b;
    return CLS_SEP;
  }
  if (b == 0xE2 && i + 2 < len && t[i + 1] == 0x84 && t[i + 2] == 0xAA) return 'k';
  if (b == 0x84 && i >= 1 && i + 1 < len && t[i - 1] == 0xE2 && t[i + 1] == 0xAA) return CLS_SKIP;
  if (b == 0xAA && i >= 2 && t[i - 2] == 0xE2 && t[i - 1] == 0x84) return CLS_SKIP;
  if (b == 0xC4 && i + 1 < len && t[i + 1] == 0xB0) return CLS_I_THEN_SEP;
  return CLS_SEP;
}

__global__ void __launch_bounds__(256) lexicon_kernel(const uint8_t *texts, const unsigned long long *offsets,
                                                      unsigned long long n_posts, double *polarity,
                                                      uint8_t *speculative, uint32_t *bull_hits, uint32_t *bear_hits) {
  const int lane = threadIdx.x & 31;
  const unsigned long long warp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned long long n_warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  for (unsigned long long post = warp; post < n_posts; post += n_warps) {
    const uint8_t *t = texts + offsets[post];
    const long long len = (long long)(offsets[post + 1] - offsets[post]);
    uint32_t bull = 0, bear = 0, spec = 0;
    for (long long i = lane; i < len; i += 32) {
      const int c = classify(t, i, len);
      if (c == CLS_SEP || c == CLS_SKIP) continue;
      // token start? look back over transparent bytes to the previous real position
      long long j = i - 1;
      while (j >= 0 && classify(t, j, len) == CLS_SKIP) --j;
      if (j >= 0) {
        const int pc = classify(t, j, len);
        if (pc > 0) continue;  // previous char continues the same token (U+0130's 'i' ends one: pc < 0)
      }
      // gather the token: up to 16 bytes, anything longer cannot match a list word
      unsigned long long lo = 0, hi = 0;
      int n = 0;
      long long p = i;
      int cc = c;
      for (;;) {
        const int ch = cc == CLS_I_THEN_SEP ? 'i' : cc;
        if (n < 8) lo |= (unsigned long long)ch << (8 * n);
        else if (n < 16) hi |= (unsigned long long)ch << (8 * (n - 8));
        ++n;
        if (cc == CLS_I_THEN_SEP || n > 16) break;
        ++p;
        while (p < len && (cc = classify(t, p, len)) == CLS_SKIP) ++p;
        if (p >= len || cc == CLS_SEP) break;
      }
      const bool too_long = n > 16;
      if (too_long) continue;
#pragma unroll
      for (int w = 0; w < N_BULL; ++w) bull += (c_bull[w].lo == lo && c_bull[w].hi == hi);
#pragma unroll
      for (int w = 0; w < N_BEAR; ++w) bear += (c_bear[w].lo == lo && c_bear[w].hi == hi);
#pragma unroll
      for (int w = 0; w < N_JARGON; ++w) spec |= (c_jargon[w].lo == lo && c_jargon[w].hi == hi);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      bull += __shfl_xor_sync(0xFFFFFFFFu, bull, o);
      bear += __shfl_xor_sync(0xFFFFFFFFu, bear, o);
      spec |= __shfl_xor_sync(0xFFFFFFFFu, spec, o);
    }
    if (lane == 0) {
      const double bh = (double)bull, eh = (double)bear;
      double pol = (bh + eh == 0.0) ? 0.0 : (bh - eh) / (bh + eh);
      if (pol != pol) pol = 0.0;
      pol = pol > 1.0 ? 1.0 : (pol < -1.0 ? -1.0 : pol);
      polarity[post] = pol;
      speculative[post] = (uint8_t)(spec != 0);
      if (bull_hits) bull_hits[post] = bull;
      if (bear_hits) bear_hits[post] = bear;
    }
  }
}

}  // namespace

void oi_set_thread_error(const std::string &msg);  // api.cu

extern "C" oi_status oi_lexicon_analyze(int32_t device, const uint8_t *texts, const uint64_t *offsets,
                                        uint64_t n_posts, double *out_polarity, uint8_t *out_speculative,
                                        uint32_t *out_bull_hits, uint32_t *out_bear_hits) {
  auto fail = [](oi_status code, const std::string &m) { oi_set_thread_error(m); return code; };
  if (n_posts == 0) return OI_OK;
  if (!offsets || !out_polarity || !out_speculative) return fail(OI_ERR_INVALID_ARG, "NULL offsets / output pointer");
  for (uint64_t i = 0; i < n_posts; ++i)
    if (offsets[i + 1] < offsets[i]) return fail(OI_ERR_INVALID_ARG, "offsets not monotone");
  const uint64_t n_bytes = offsets[n_posts] - offsets[0];
  if (n_bytes && !texts) return fail(OI_ERR_INVALID_ARG, "texts is NULL");
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0) return fail(OI_ERR_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU fallback)");
  if (device < 0 || device >= n_dev) return fail(OI_ERR_INVALID_ARG, "device ordinal out of range");
  uint8_t *d_text = nullptr, *d_spec = nullptr;
  unsigned long long *d_off = nullptr;
  double *d_pol = nullptr;
  uint32_t *d_bull = nullptr, *d_bear = nullptr;
  cudaStream_t st = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_text); cudaFree(d_spec); cudaFree(d_off); cudaFree(d_pol); cudaFree(d_bull); cudaFree(d_bear);
    if (st) cudaStreamDestroy(st);
  };
#define LX_CK(call)                                                                                  \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) {                                                                         \
      cleanup();                                                                                     \
      return fail(e_ == cudaErrorMemoryAllocation ? OI_ERR_OUT_OF_MEMORY : OI_ERR_CUDA,              \
                  std::string(#call " failed: ") + cudaGetErrorString(e_));                          \
    }                                                                                                \
  } while (0)
  LX_CK(cudaSetDevice(device));
  LX_CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  LX_CK(cudaMalloc(&d_text, n_bytes ? n_bytes : 1));
  LX_CK(cudaMalloc(&d_off, (n_posts + 1) * sizeof(unsigned long long)));
  LX_CK(cudaMalloc(&d_pol, n_posts * sizeof(double)));
  LX_CK(cudaMalloc(&d_spec, n_posts));
  LX_CK(cudaMalloc(&d_bull, n_posts * sizeof(uint32_t)));
  LX_CK(cudaMalloc(&d_bear, n_posts * sizeof(uint32_t)));
  if (n_bytes) LX_CK(cudaMemcpyAsync(d_text, texts + offsets[0], n_bytes, cudaMemcpyHostToDevice, st));
  // offsets are rebased to the copied blob on the device side by passing texts - offsets[0]
  LX_CK(cudaMemcpyAsync(d_off, offsets, (n_posts + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
  uint64_t blocks = (n_posts + 7) / 8;
  if (blocks > 148ull * 16) blocks = 148ull * 16;
  lexicon_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_text - offsets[0], d_off, n_posts, d_pol, d_spec, d_bull, d_bear);
  LX_CK(cudaGetLastError());
  LX_CK(cudaMemcpyAsync(out_polarity, d_pol, n_posts * sizeof(double), cudaMemcpyDeviceToHost, st));
  LX_CK(cudaMemcpyAsync(out_speculative, d_spec, n_posts, cudaMemcpyDeviceToHost, st));
  if (out_bull_hits) LX_CK(cudaMemcpyAsync(out_bull_hits, d_bull, n_posts * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  if (out_bear_hits) LX_CK(cudaMemcpyAsync(out_bear_hits, d_bear, n_posts * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  LX_CK(cudaStreamSynchronize(st));
#undef LX_CK
  cleanup();
  return OI_OK;
}
