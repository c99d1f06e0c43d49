// Deterministic synthetic workloads: Zipfian corpora and query sets.
// Language-neutral PRNG (SplitMix64) so any binding regenerates the same data.
//   uniform double = (x >> 11) * 2^-53
//   vocabulary     = V distinct lowercase strings over a..z, length uniform 3..10, generated in
//                    rank order (duplicates rejected); P(rank r) ~ 1/r  (Zipf s = 1.0) through an
//                    inverse-CDF table + binary search
//   documents      = per field a token count uniform in [len_min, len_max], tokens Zipf-drawn;
//                    every document has its own PRNG stream (seed, doc index) so the corpus does
//                    not depend on how generation is chunked
// This library only GENERATES inputs; it is used by tests and bench.py for both the CUDA path
// and the CPU oracle.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_set>
#include <vector>

namespace {

struct SplitMix64 {
  uint64_t s;
  explicit SplitMix64(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
  }
  double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  uint64_t below(uint64_t n) { return (uint64_t)(uniform() * (double)n); }   // n << 2^53
};

// Seed of document d's private stream: the doc index goes through the SplitMix64 finaliser first.
// (Seeding with seed + gamma * d would make doc d + 1's stream doc d's stream shifted by one draw:
// neighbouring documents would share all but one token.)
static inline uint64_t doc_stream_seed(uint64_t seed, uint64_t d) {
  SplitMix64 h(seed ^ (0xD1B54A32D192ED03ULL * (d + 1)));
  return h.next();
}

struct Workload {
  uint64_t seed;
  uint32_t V, F;
  uint32_t len_min[4], len_max[4];
  std::vector<std::string> vocab;
  std::vector<double> cdf;

  uint32_t zipf(SplitMix64& r) const {
    double u = r.uniform();
    return (uint32_t)(std::upper_bound(cdf.begin(), cdf.end(), u) - cdf.begin());
  }
};

}  // namespace

extern "C" {

void* wl_new(uint64_t seed, uint32_t vocab_size, uint32_t n_fields, const uint32_t* len_min, const uint32_t* len_max) {
  Workload* w = new Workload();
  w->seed = seed; w->V = vocab_size; w->F = n_fields;
  for (uint32_t f = 0; f < n_fields && f < 4; ++f) { w->len_min[f] = len_min[f]; w->len_max[f] = len_max[f]; }
  SplitMix64 r(seed ^ 0x766F636162ULL);   // "vocab"
  std::unordered_set<std::string> seen;
  w->vocab.reserve(vocab_size);
  while (w->vocab.size() < vocab_size) {
    uint32_t len = 3 + (uint32_t)r.below(8);
    std::string s(len, 'a');
    for (uint32_t i = 0; i < len; ++i) s[i] = (char)('a' + r.below(26));
    if (seen.insert(s).second) w->vocab.push_back(std::move(s));
  }
  w->cdf.resize(vocab_size);
  double h = 0.0;
  for (uint32_t i = 0; i < vocab_size; ++i) h += 1.0 / (double)(i + 1);
  double acc = 0.0;
  for (uint32_t i = 0; i < vocab_size; ++i) { acc += (1.0 / (double)(i + 1)) / h; w->cdf[i] = acc; }
  w->cdf.back() = 2.0;   // guard: upper_bound never runs off the end
  return w;
}

void wl_free(void* h) { delete (Workload*)h; }

uint32_t wl_vocab_word(void* h, uint32_t rank, uint8_t* out, uint32_t cap) {
  Workload* w = (Workload*)h;
  const std::string& s = w->vocab[rank];
  std::memcpy(out, s.data(), std::min<size_t>(cap, s.size()));
  return (uint32_t)s.size();
}

// Upper bounds for the buffers of wl_gen_docs over n docs.
void wl_doc_bounds(void* h, uint64_t n_docs, uint64_t* max_tokens, uint64_t* max_bytes) {
  Workload* w = (Workload*)h;
  uint64_t t = 0;
  for (uint32_t f = 0; f < w->F; ++f) t += w->len_max[f];
  *max_tokens = t * n_docs;
  *max_bytes = t * n_docs * 10;
}

// Documents [d0, d1): tokens concatenated in (doc, field, token) order.
void wl_gen_docs(void* h, uint64_t d0, uint64_t d1, uint8_t* tok_bytes, uint64_t* tok_off,
                 uint32_t* field_tok_count, uint64_t* n_tokens, uint64_t* n_bytes) {
  Workload* w = (Workload*)h;
  uint64_t t = 0, b = 0;
  tok_off[0] = 0;
  for (uint64_t d = d0; d < d1; ++d) {
    SplitMix64 r(doc_stream_seed(w->seed, d));
    r.next();
    for (uint32_t f = 0; f < w->F; ++f) {
      uint32_t n = w->len_min[f] + (uint32_t)r.below(w->len_max[f] - w->len_min[f] + 1);
      field_tok_count[(d - d0) * w->F + f] = n;
      for (uint32_t i = 0; i < n; ++i) {
        const std::string& s = w->vocab[w->zipf(r)];
        std::memcpy(tok_bytes + b, s.data(), s.size());
        b += s.size();
        tok_off[++t] = b;
      }
    }
  }
  *n_tokens = t;
  *n_bytes = b;
}

// Queries.  mode 0: one full vocabulary word, Zipf-drawn (cfg 1/3/4).
//           mode 1: 2..4 terms, each a Zipf-drawn word truncated to its first 2..4 chars (cfg 2).
//           mode 2: one term = the first 1..2 chars of a Zipf-drawn word (cfg 0 bench-shape corpus).
void wl_gen_queries(void* h, uint64_t qseed, uint64_t n_queries, uint32_t mode, uint64_t* query_term_off,
                    uint64_t* term_byte_off, uint8_t* term_bytes, uint64_t* n_terms, uint64_t* n_bytes) {
  Workload* w = (Workload*)h;
  SplitMix64 r(qseed);
  uint64_t t = 0, b = 0;
  query_term_off[0] = 0;
  term_byte_off[0] = 0;
  for (uint64_t q = 0; q < n_queries; ++q) {
    uint32_t nt = (mode == 1) ? 2 + (uint32_t)r.below(3) : 1;
    for (uint32_t i = 0; i < nt; ++i) {
      const std::string& s = w->vocab[w->zipf(r)];
      size_t len = s.size();
      if (mode == 1) len = std::min<size_t>(len, 2 + r.below(3));
      else if (mode == 2) len = std::min<size_t>(len, 1 + r.below(2));
      std::memcpy(term_bytes + b, s.data(), len);
      b += len;
      term_byte_off[++t] = b;
    }
    query_term_off[q + 1] = t;
  }
  *n_terms = t;
  *n_bytes = b;
}

// cfg 0 "bench shape" corpus (benches/test_benchmark.rs:21-43): title = two random 5-letter words
// over the 24-char string "abcdefghilkjapqrstuvwxyz", one field.
void wl_gen_bench_docs(uint64_t seed, uint64_t d0, uint64_t d1, uint8_t* tok_bytes, uint64_t* tok_off,
                       uint32_t* field_tok_count, uint64_t* n_tokens, uint64_t* n_bytes) {
  static const char kAllowed[] = "abcdefghilkjapqrstuvwxyz";
  uint64_t t = 0, b = 0;
  tok_off[0] = 0;
  for (uint64_t d = d0; d < d1; ++d) {
    SplitMix64 r(doc_stream_seed(seed, d));
    r.next();
    field_tok_count[d - d0] = 2;
    for (int wd = 0; wd < 2; ++wd) {
      for (int i = 0; i < 5; ++i) tok_bytes[b++] = (uint8_t)kAllowed[r.below(24)];
      tok_off[++t] = b;
    }
  }
  *n_tokens = t;
  *n_bytes = b;
}

// ordinals of the documents "removed" in cfg 4: a fraction of [0, n_docs) drawn without replacement
uint64_t wl_gen_removed(uint64_t seed, uint64_t n_docs, double fraction, uint64_t* out) {
  SplitMix64 r(seed ^ 0x72656D6F7665ULL);   // "remove"
  uint64_t n = 0;
  for (uint64_t d = 0; d < n_docs; ++d)
    if (r.uniform() < fraction) out[n++] = d;
  return n;
}

}  // extern "C"
