// probly_b200.hpp — header-only C++17 mirror of the reference's public surface for the query
// path, over the C ABI of probly_b200.h.  Same names, argument meaning and error behaviour as
// probly-search 2.0.1 where C++ allows:
//   probly::Index<T>{fields_num}            Index::<T>::new              src/index.rs:37
//   add_document(accessors, tokenizer, key, doc)                         src/index.rs:77
//   remove_document(key) / vacuum()                                      src/index.rs:161 / :194
//   query(query, calculator, tokenizer, fields_boost) -> vector<QueryResult<T>>   src/query.rs:21
//   probly::score::bm25::make() / zero_to_one::make()   (`new` is a C++ keyword)  bm25.rs:21, zero_to_one.rs:35
// The reference panics on its error paths (unwrap(), src/query.rs:46,63,70); this mirror throws
// probly::Error.  Only the two built-in calculators can run on the device; anything else does
// not compile (no CPU fallback).
#pragma once
#include <algorithm>
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <string_view>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "probly_b200.h"

namespace probly {

struct Error : std::runtime_error {
  int code;
  Error(int c, const char* msg) : std::runtime_error(msg), code(c) {}
};
inline void check(int rc) {
  if (rc != PB_OK) throw Error(rc, pb_last_error());
}

template <class T>
struct QueryResult {   // src/query.rs:9-15
  T key;
  double score;
  bool operator==(const QueryResult& o) const { return key == o.key && score == o.score; }
};

using Tokenizer = std::vector<std::string> (*)(std::string_view);                 // src/lib.rs:14
template <class D> using FieldAccessor = std::vector<std::string_view> (*)(const D&);   // src/lib.rs:11

namespace score {
namespace bm25 {
struct BM25 { double bm25k1 = 1.2, bm25b = 0.75; };   // bm25.rs:14-26
inline BM25 make() { return BM25{}; }
}  // namespace bm25
namespace zero_to_one {
struct ZeroToOne {};                                    // zero_to_one.rs:24-39
inline ZeroToOne make() { return ZeroToOne{}; }
}  // namespace zero_to_one
template <class S> struct device_scorer : std::false_type {};
template <> struct device_scorer<bm25::BM25> : std::true_type {};
template <> struct device_scorer<zero_to_one::ZeroToOne> : std::true_type {};
}  // namespace score

template <class T, class Hash = std::hash<T>>
class Index {
 public:
  explicit Index(size_t fields_num, int device = 0) : fields_num_(fields_num), device_(device) {
    check(pb_builder_create((uint32_t)fields_num, &b_));
  }
  ~Index() {
    if (ix_) pb_index_destroy(ix_);
    if (b_) pb_builder_destroy(b_);
  }
  Index(const Index&) = delete;
  Index& operator=(const Index&) = delete;

  template <class D>
  void add_document(const std::vector<FieldAccessor<D>>& field_accessors, Tokenizer tokenizer, T key, const D& doc) {
    std::string bytes;
    std::vector<uint64_t> off{0};
    std::vector<uint32_t> vcount, fcount;
    for (size_t i = 0; i < fields_num_; ++i) {
      auto values = field_accessors[i](doc);
      fcount.push_back((uint32_t)values.size());
      for (auto v : values) {
        auto terms = tokenizer(v);
        vcount.push_back((uint32_t)terms.size());
        for (auto& t : terms) { bytes += t; off.push_back(bytes.size()); }
      }
    }
    vcount.push_back(0);
    pb_doc_tokens d{(const uint8_t*)bytes.data(), off.data(), vcount.data(), fcount.data()};
    check(pb_builder_add_document(b_, key_id(key), &d));
    dirty_ = true;
  }
  void remove_document(T key) {
    auto it = key_to_id_.find(key);
    if (it == key_to_id_.end()) return;
    check(pb_builder_remove_document(b_, it->second));
    dirty_ = true;
  }
  void vacuum() { check(pb_builder_vacuum(b_)); dirty_ = true; }

  // Not in the reference (it has no serialisation): writes the flattened image a serving process can
  // load with pb_image_load + pb_index_create, without building this mutable index.
  void save_image(const std::string& path) {
    pb_index_image im;
    check(pb_builder_flatten(b_, &im));
    check(pb_image_save(&im, path.c_str()));
  }

  template <class S>
  std::vector<QueryResult<T>> query(std::string_view query, S& score_calculator, Tokenizer tokenizer,
                                    const std::vector<double>& fields_boost) {
    static_assert(score::device_scorer<S>::value,
                  "only score::bm25::BM25 and score::zero_to_one::ZeroToOne exist as device code; there is no CPU fallback");
    sync_device();
    auto terms = tokenizer(query);
    std::string bytes;
    std::vector<uint64_t> toff{0};
    for (auto& t : terms) { bytes += t; toff.push_back(bytes.size()); }
    uint64_t qoff[2] = {0, terms.size()};
    pb_query_batch_desc d{};
    d.n_queries = 1; d.query_term_off = qoff; d.term_byte_off = toff.data(); d.term_bytes = (const uint8_t*)bytes.data();
    if constexpr (std::is_same_v<S, score::bm25::BM25>) { d.scorer = PB_SCORER_BM25; d.bm25_k1 = score_calculator.bm25k1; d.bm25_b = score_calculator.bm25b; }
    else { d.scorer = PB_SCORER_ZERO_TO_ONE; d.bm25_k1 = 1.2; d.bm25_b = 0.75; }
    d.fields_boost = fields_boost.data(); d.n_fields_boost = (uint32_t)fields_boost.size(); d.top_k = 0;
    uint64_t cap = 1024, n = 0;
    std::vector<uint32_t> oq, od;
    std::vector<double> os;
    for (;;) {
      oq.assign(cap, 0); od.assign(cap, 0); os.assign(cap, 0.0);
      int rc = pb_query_full(ix_, &d, cap, oq.data(), od.data(), os.data(), &n);
      if (rc == PB_ERR_CAPACITY) { cap = n + 16; continue; }
      check(rc);
      break;
    }
    std::vector<QueryResult<T>> res;
    for (uint64_t i = 0; i < n; ++i) res.push_back({id_to_key_[ord_to_id_[od[i]]], os[i]});
    // src/query.rs:103 sorts by score desc; exactly tied scores by ordinal here (hash order there)
    std::vector<size_t> idx(n);
    for (size_t i = 0; i < n; ++i) idx[i] = i;
    std::sort(idx.begin(), idx.end(), [&](size_t a, size_t b2) { return os[a] != os[b2] ? os[a] > os[b2] : od[a] < od[b2]; });
    std::vector<QueryResult<T>> sorted;
    for (size_t i : idx) sorted.push_back(res[i]);
    return sorted;
  }

 private:
  uint64_t key_id(const T& key) {
    auto it = key_to_id_.find(key);
    if (it != key_to_id_.end()) return it->second;
    uint64_t id = id_to_key_.size();
    key_to_id_.emplace(key, id);
    id_to_key_.push_back(key);
    return id;
  }
  void sync_device() {
    if (ix_ && !dirty_) return;
    pb_index_image im{};
    check(pb_builder_flatten(b_, &im));
    if (ix_) { pb_index_destroy(ix_); ix_ = nullptr; }
    check(pb_index_create(&im, device_, &ix_));
    ord_to_id_.assign(im.doc_key, im.doc_key + im.n_docs);
    dirty_ = false;
  }
  size_t fields_num_;
  int device_;
  pb_builder* b_ = nullptr;
  pb_index* ix_ = nullptr;
  bool dirty_ = true;
  std::unordered_map<T, uint64_t, Hash> key_to_id_;
  std::vector<T> id_to_key_;
  std::vector<uint64_t> ord_to_id_;
};

}  // namespace probly
