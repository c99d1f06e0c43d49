// =============================================================================
//  ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.
//
//  A structure-faithful CPU restatement (C++17) of the probly-search query hot
//  path and of the index construction that defines the data it reads.  It keeps
//  the reference's *memory behaviour* on purpose (generational slot arenas,
//  prepended intrusive linked lists, one posting node per term OCCURRENCE, hash
//  map for docs / scores, hash sets for removed / visited, a second full walk
//  for count_documents, a root re-descent per expansion) because it doubles as
//  the timed CPU baseline (`bench.py` cpu_baseline leg, kind = "port").
//
//  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
//  --impl reference legs may load this library.  The product library
//  (probly_search_b200/csrc) never links or calls it.
//
//  PARITY STATUS: pinned.  tests/test_oracle_goldens.py checks this file against
//  every golden vector the reference's own tests hold for the path
//  (SURVEY.md §8c): src/score/default/bm25.rs:104-136, src/query.rs:181-387,
//  src/score/default/zero_to_one.rs:138-404, tests/integrations_tests.rs:27-149,
//  tests/document_frequency.rs:5-32, src/index.rs:492-784.
//  NOT pinned by any reference test: BM25 with boosts != 1, the
//  pre-vacuum removed mask under BM25, the merger across several terms AND
//  expansions, non-ASCII byte lengths, multi-valued fields.  There this
//  restatement of the cited lines is the only authority.
//  The real crate cannot be compiled here (no rustc/cargo, un-vendored
//  hashbrown 0.14 / typed-generational-arena 0.2), so there is no oracle/_ref.
//
//  Each function cites the reference file:line it follows (paths relative to
//  the reference root).
// =============================================================================
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

namespace orc {

// ----------------------------------------------------------------------------
// Containers standing in for the two third-party crates (containers only; no
// arithmetic lives in them — SURVEY.md §2 #17).
// ----------------------------------------------------------------------------

// typed-generational-arena 0.2 `StandardIndex`: (slot, non-zero generation).
// gen == 0 plays the role of Option::None (the crate uses a NonZero niche).
struct Aidx {
  uint64_t slot = 0;
  uint64_t gen = 0;
  bool some() const { return gen != 0; }
  bool operator==(const Aidx& o) const { return slot == o.slot && gen == o.gen; }
};

// `StandardArena<V>`: slot vector + free list + generation check on access.
template <class V>
class Arena {
  struct Slot {
    V value;
    uint64_t gen;   // 0 = free
    uint64_t next_free;
  };
  std::vector<Slot> slots_;
  uint64_t free_head_ = UINT64_MAX;
  uint64_t generation_ = 1;
  size_t len_ = 0;

 public:
  void reserve(size_t n) { slots_.reserve(n); }
  Aidx insert(V&& v) {
    ++len_;
    if (free_head_ != UINT64_MAX) {
      uint64_t s = free_head_;
      free_head_ = slots_[s].next_free;
      slots_[s].value = std::move(v);
      slots_[s].gen = generation_;
      return Aidx{s, generation_};
    }
    slots_.push_back(Slot{std::move(v), generation_, UINT64_MAX});
    return Aidx{slots_.size() - 1, generation_};
  }
  V* get(Aidx i) {
    if (i.slot >= slots_.size()) return nullptr;
    Slot& s = slots_[i.slot];
    return (s.gen == i.gen && i.gen != 0) ? &s.value : nullptr;
  }
  const V* get(Aidx i) const { return const_cast<Arena*>(this)->get(i); }
  void remove(Aidx i) {
    if (!get(i)) return;
    Slot& s = slots_[i.slot];
    s.value = V{};
    s.gen = 0;
    s.next_free = free_head_;
    free_head_ = i.slot;
    ++generation_;
    --len_;
  }
  bool is_empty() const { return len_ == 0; }
  size_t len() const { return len_; }
};

static inline uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
  x ^= x >> 27; x *= 0x94d049bb133111ebULL;
  x ^= x >> 31;
  return x;
}

// hashbrown::HashMap<usize, V> stand-in: open addressing, linear probing,
// grows from empty by doubling (so the rehash-on-growth cost of the reference's
// `HashMap::new()` per query is kept), tombstone-free erase by backward shift.
template <class V>
class FlatMap {
  struct Cell { uint64_t key; V val; bool used; };
  std::vector<Cell> cells_;
  size_t len_ = 0;
  size_t mask_ = 0;

  void grow() {
    size_t ncap = cells_.empty() ? 4 : cells_.size() * 2;
    std::vector<Cell> old;
    old.swap(cells_);
    cells_.resize(ncap);
    for (auto& c : cells_) c.used = false;
    mask_ = ncap - 1;
    len_ = 0;
    for (auto& c : old) if (c.used) *insert_slot(c.key) = std::move(c.val);
  }
  V* insert_slot(uint64_t key) {
    size_t i = mix64(key) & mask_;
    while (cells_[i].used) {
      if (cells_[i].key == key) return &cells_[i].val;
      i = (i + 1) & mask_;
    }
    cells_[i].used = true;
    cells_[i].key = key;
    ++len_;
    return &cells_[i].val;
  }

 public:
  size_t len() const { return len_; }
  void clear() { cells_.clear(); len_ = 0; mask_ = 0; }
  V* get(uint64_t key) {
    if (cells_.empty()) return nullptr;
    size_t i = mix64(key) & mask_;
    while (cells_[i].used) {
      if (cells_[i].key == key) return &cells_[i].val;
      i = (i + 1) & mask_;
    }
    return nullptr;
  }
  const V* get(uint64_t key) const { return const_cast<FlatMap*>(this)->get(key); }
  bool contains(uint64_t key) const { return get(key) != nullptr; }
  // insert-or-overwrite, like HashMap::insert
  void insert(uint64_t key, V v) {
    if ((len_ + 1) * 8 > cells_.size() * 7) grow();   // 7/8 load factor, as SwissTable
    *insert_slot(key) = std::move(v);
  }
  bool remove(uint64_t key) {
    if (cells_.empty()) return false;
    size_t i = mix64(key) & mask_;
    while (cells_[i].used && cells_[i].key != key) i = (i + 1) & mask_;
    if (!cells_[i].used) return false;
    size_t j = i;
    for (;;) {
      j = (j + 1) & mask_;
      if (!cells_[j].used) break;
      size_t h = mix64(cells_[j].key) & mask_;
      bool between = (i <= j) ? (i < h && h <= j) : (i < h || h <= j);
      if (!between) { cells_[i] = std::move(cells_[j]); i = j; }
    }
    cells_[i].used = false;
    cells_[i].val = V{};
    --len_;
    return true;
  }
  template <class Fn> void for_each(Fn&& fn) {
    for (auto& c : cells_) if (c.used) fn(c.key, c.val);
  }
};
struct Unit {};
using FlatSet = FlatMap<Unit>;

// Rust `str::chars()`: UTF-8 → Unicode scalar values.  Inputs are valid UTF-8
// (Rust &str guarantees it; the Python wrapper encodes str → UTF-8).
static inline uint32_t next_char(const char*& p, const char* end) {
  uint8_t c = (uint8_t)*p++;
  if (c < 0x80) return c;
  int extra = (c >= 0xF0) ? 3 : (c >= 0xE0) ? 2 : 1;
  uint32_t cp = c & (0x3F >> extra);
  while (extra-- > 0 && p < end) cp = (cp << 6) | ((uint8_t)*p++ & 0x3F);
  return cp;
}

// ----------------------------------------------------------------------------
// src/index.rs data types
// ----------------------------------------------------------------------------
struct InvertedIndexNode {      // index.rs:364-373
  uint32_t ch = 0;
  Aidx next, first_child, first_doc;
};
struct DocumentPointer {        // index.rs:354-361
  Aidx next;
  uint64_t details_key = 0;
  std::vector<size_t> term_frequency;
};
struct DocumentDetails {        // index.rs:342-349
  uint64_t key = 0;
  std::vector<size_t> field_length;
};
struct FieldDetails {           // index.rs:391-396
  size_t sum = 0;
  double avg = 0.0;
};

struct QueryResult { uint64_t key; double score; };   // query.rs:9-15

// instrumentation: number of ScoreCalculator::score calls (query.rs:67) made by this thread
static thread_local uint64_t tl_score_calls = 0;

struct TermData {               // score/calculator.rs:9-19
  size_t query_term_index;
  std::string_view query_term;
  std::string_view query_term_expanded;
  size_t query_terms_len;
};
struct FieldData {              // score/calculator.rs:21-26
  const double* fields_boost;
  const FieldDetails* fields;
};

class Index {
 public:
  FlatMap<DocumentDetails> docs;                 // index.rs:21
  Aidx root;                                     // index.rs:23
  std::vector<FieldDetails> fields;              // index.rs:25
  Arena<InvertedIndexNode> arena_index;          // index.rs:27
  Arena<DocumentPointer> arena_doc;              // index.rs:28
  bool removed_some = false;                     // index.rs:32  Option<HashSet<T>>
  FlatSet removed;

  // index.rs:37-60
  explicit Index(size_t fields_num, size_t expected_index_size = 1000,
                 size_t expected_documents_count = 10000) {
    fields.assign(fields_num, FieldDetails{});
    arena_index.reserve(expected_index_size);
    arena_doc.reserve(expected_documents_count);
    InvertedIndexNode r;
    r.ch = 0;
    root = arena_index.insert(std::move(r));
  }

  // index.rs:321-337 — linear scan of the sibling list
  Aidx find_child_by_char(const InvertedIndexNode* from, uint32_t ch) const {
    Aidx it = from->first_child;
    while (it.some()) {
      const InvertedIndexNode* n = arena_index.get(it);
      if (n->ch == ch) return it;
      it = n->next;
    }
    return Aidx{};
  }

  // index.rs:300-318
  Aidx find_node(Aidx from, std::string_view term) const {
    Aidx it = from;
    bool have = true;
    const char* p = term.data();
    const char* e = p + term.size();
    while (p < e) {
      uint32_t ch = next_char(p, e);
      if (have) {
        Aidx nx = find_child_by_char(arena_index.get(it), ch);
        if (nx.some()) it = nx; else have = false;
      } else {
        break;
      }
    }
    return have ? it : Aidx{};
  }

  // index.rs:409-419 — children are PREPENDED
  void add_child(Aidx parent, Aidx child) {
    InvertedIndexNode* p = arena_index.get(parent);
    if (p->first_child.some()) arena_index.get(child)->next = p->first_child;
    arena_index.get(parent)->first_child = child;
  }

  // index.rs:437-452
  Aidx create_nodes(Aidx parent, std::string_view term, size_t start) {
    const char* p = term.data();
    const char* e = p + term.size();
    size_t i = 0;
    while (p < e) {
      uint32_t ch = next_char(p, e);
      if (i++ < start) continue;
      InvertedIndexNode n;
      n.ch = ch;
      Aidx nn = arena_index.insert(std::move(n));
      add_child(parent, nn);
      parent = arena_index.get(parent)->first_child;
    }
    return parent;
  }

  // index.rs:77-158.  `field_values[i]` = the values field accessor i returned,
  // each already run through the tokenizer (empty tokens still present).
  void add_document(uint64_t key,
                    const std::vector<std::vector<std::vector<std::string_view>>>& field_values) {
    std::vector<size_t> field_length(fields.size(), 0);
    std::unordered_map<std::string_view, std::vector<size_t>> term_counts;
    std::vector<std::string_view> all_terms;
    const size_t fields_len = fields.size();
    for (size_t i = 0; i < fields_len; ++i) {
      FieldDetails& fd = fields[i];
      for (const auto& terms : field_values[i]) {
        size_t filtered = 0;                                   // index.rs:99
        for (std::string_view term : terms) {
          if (!term.empty()) {
            ++filtered;
            all_terms.push_back(term);                         // one entry per OCCURRENCE (index.rs:103)
            auto it = term_counts.find(term);
            if (it == term_counts.end())
              it = term_counts.emplace(term, std::vector<size_t>(fields_len, 0)).first;
            it->second[i] += 1;
          }
        }
        fd.sum += filtered;                                    // index.rs:112
        fd.avg = (double)fd.sum / ((double)docs.len() + 1.0);  // index.rs:113 (before the doc is inserted)
        field_length[i] = filtered;                            // index.rs:114 (last value wins)
      }
    }
    docs.insert(key, DocumentDetails{key, field_length});      // index.rs:118
    for (std::string_view term : all_terms) {                  // index.rs:119-157
      Aidx node_index = root;
      const char* p = term.data();
      const char* e = p + term.size();
      size_t i = 0;
      while (p < e) {
        uint32_t ch = next_char(p, e);
        const InvertedIndexNode* node = arena_index.get(node_index);
        if (!node->first_child.some()) {
          node_index = create_nodes(node_index, term, i);
          break;
        }
        Aidx nx = find_child_by_char(node, ch);
        if (!nx.some()) {
          node_index = create_nodes(node_index, term, i);
          break;
        }
        node_index = nx;
        ++i;
      }
      DocumentPointer dp;
      dp.details_key = key;
      dp.term_frequency = term_counts[term];                   // clone per occurrence (index.rs:153)
      // index.rs:422-433 — postings are PREPENDED
      InvertedIndexNode* n = arena_index.get(node_index);
      if (n->first_doc.some()) dp.next = n->first_doc;
      Aidx di = arena_doc.insert(std::move(dp));
      arena_index.get(node_index)->first_doc = di;
    }
  }

  // index.rs:161-191
  void remove_document(uint64_t key) {
    removed_some = true;
    const DocumentDetails* d = docs.get(key);
    bool remove_key = false;
    if (d) {
      removed.insert(key, Unit{});
      remove_key = true;
      double new_len = (double)(docs.len() - 1);
      for (size_t i = 0; i < fields.size(); ++i) {
        size_t fl = d->field_length[i];
        if (fl > 0) {
          fields[i].sum -= fl;
          fields[i].avg = (double)fields[i].sum / new_len;
        }
      }
    }
    if (remove_key) docs.remove(key);
  }

  // index.rs:245-279
  size_t disconnect_and_count_documents(Aidx node_index, const FlatSet* rem) {
    InvertedIndexNode* node = arena_index.get(node_index);
    Aidx prev{};
    Aidx ptr = node->first_doc;
    size_t df = 0;
    while (ptr.some()) {
      DocumentPointer* dp = arena_doc.get(ptr);
      bool is_removed = rem && rem->contains(dp->details_key);
      Aidx nxt = dp->next;
      if (is_removed) {
        if (!prev.some()) node->first_doc = nxt;
        else arena_doc.get(prev)->next = nxt;
      } else {
        ++df;
        prev = ptr;
      }
      if (is_removed) arena_doc.remove(ptr);
      ptr = nxt;
    }
    return df;
  }

  // index.rs:202-241
  size_t vacuum_node(Aidx node_index, const FlatSet* rem) {
    disconnect_and_count_documents(node_index, rem);
    Aidx prev_child{};
    size_t ret = 0;
    const InvertedIndexNode* node = arena_index.get(node_index);
    if (node->first_doc.some()) ret = 1;
    Aidx child = node->first_child;
    while (child.some()) {
      size_t r = vacuum_node(child, rem);
      ret |= r;
      Aidx child_next = arena_index.get(child)->next;
      if (r == 0) {
        if (prev_child.some()) arena_index.get(prev_child)->next = child_next;
        else arena_index.get(node_index)->first_child = child_next;
      } else {
        prev_child = child;
      }
      if (r == 0) arena_index.remove(child);
      child = child_next;
    }
    return ret;
  }

  // index.rs:194-199
  void vacuum() {
    FlatSet rem;
    std::swap(rem, removed);
    vacuum_node(root, &rem);
    removed.clear();
    removed_some = false;
  }

  // index.rs:282-297 — a FULL extra walk of the list
  size_t count_documents(Aidx node_index) const {
    const InvertedIndexNode* node = arena_index.get(node_index);
    Aidx ptr = node->first_doc;
    size_t df = 0;
    while (ptr.some()) {
      const DocumentPointer* dp = arena_doc.get(ptr);
      bool is_removed = removed_some ? removed.contains(dp->details_key) : false;
      if (!is_removed) ++df;
      ptr = dp->next;
    }
    return df;
  }

  // query.rs:130-147 — recursive DFS; a fresh String per edge
  void expand_term_from_node(const InvertedIndexNode* node, std::vector<std::string>& results,
                             const std::string& term) const {
    if (node->first_doc.some()) results.push_back(term);
    Aidx child = node->first_child;
    while (child.some()) {
      const InvertedIndexNode* cb = arena_index.get(child);
      std::string inter = term;
      // push the child's char back as UTF-8
      uint32_t c = cb->ch;
      if (c < 0x80) inter.push_back((char)c);
      else if (c < 0x800) { inter.push_back((char)(0xC0 | (c >> 6))); inter.push_back((char)(0x80 | (c & 0x3F))); }
      else if (c < 0x10000) { inter.push_back((char)(0xE0 | (c >> 12))); inter.push_back((char)(0x80 | ((c >> 6) & 0x3F))); inter.push_back((char)(0x80 | (c & 0x3F))); }
      else { inter.push_back((char)(0xF0 | (c >> 18))); inter.push_back((char)(0x80 | ((c >> 12) & 0x3F))); inter.push_back((char)(0x80 | ((c >> 6) & 0x3F))); inter.push_back((char)(0x80 | (c & 0x3F))); }
      expand_term_from_node(cb, results, inter);
      child = cb->next;
    }
  }

  // query.rs:109-126
  std::vector<std::string> expand_term(std::string_view term) const {
    Aidx node = find_node(root, term);
    std::vector<std::string> results;
    if (node.some()) expand_term_from_node(arena_index.get(node), results, std::string(term));
    return results;
  }

  // query.rs:21-106
  template <class S>
  std::vector<QueryResult> query(const std::vector<std::string_view>& query_terms, S& calc,
                                 const double* fields_boost) {
    FlatMap<double> scores;                                            // query.rs:31
    const size_t query_terms_len = query_terms.size();                 // query.rs:32 (counts "" tokens)
    for (size_t qti = 0; qti < query_terms.size(); ++qti) {
      std::string_view query_term = query_terms[qti];
      if (query_term.empty()) continue;                                // query.rs:35
      std::vector<std::string> expanded_terms = expand_term(query_term);
      FlatSet visited;                                                 // query.rs:37 (per QUERY TERM)
      for (const std::string& expanded : expanded_terms) {
        Aidx term_node = find_node(root, expanded);                    // re-descent from the root (query.rs:39-43)
        if (!term_node.some()) continue;
        size_t df = count_documents(term_node);                        // query.rs:45
        const InvertedIndexNode* tn = arena_index.get(term_node);
        if (!tn->first_doc.some() || df == 0) continue;                // query.rs:47-48
        TermData td{qti, query_term, expanded, query_terms_len};
        auto pre = calc.before_each(td, df, docs.len());               // query.rs:55-59
        Aidx ptr = tn->first_doc;
        while (ptr.some()) {                                           // HOT LOOP query.rs:61-89
          const DocumentPointer* dp = arena_doc.get(ptr);
          uint64_t key = dp->details_key;
          if (!removed_some || !removed.contains(key)) {
            FieldData fdta{fields_boost, fields.data()};
            ++tl_score_calls;
            double s;
            bool some = calc.score(pre, *dp, *docs.get(key), term_node, fdta, td, &s);
            if (some) {
              // max_score_merger, query.rs:150-164
              const double* prev = scores.get(key);
              double ns;
              if (prev) ns = visited.contains(key) ? std::fmax(*prev, s) : (*prev + s);
              else ns = s;
              scores.insert(key, ns);
            }
          }
          visited.insert(key, Unit{});                                 // query.rs:87 (even if removed / None)
          ptr = dp->next;
        }
      }
    }
    std::vector<QueryResult> result;                                   // query.rs:97-100
    scores.for_each([&](uint64_t k, double& v) { result.push_back(QueryResult{k, v}); });
    calc.finalize(result);                                             // query.rs:101
    // query.rs:103 sorts by score desc only (ties in hash order — unspecified).  The
    // reference's own comparison rule (lib.rs:54-58) re-sorts by (score desc, key asc);
    // that canonical order is what we return.
    std::sort(result.begin(), result.end(), [](const QueryResult& a, const QueryResult& b) {
      if (a.score != b.score) return a.score > b.score;
      return a.key < b.key;
    });
    return result;
  }
};

// ----------------------------------------------------------------------------
// src/score/default/bm25.rs
// ----------------------------------------------------------------------------
struct BM25 {
  double bm25k1 = 1.2;   // bm25.rs:21-26
  double bm25b = 0.75;
  struct Pre { double idf; double expansion_boost; };

  // bm25.rs:35-58
  Pre before_each(const TermData& t, size_t document_frequency, size_t documents_len) {
    size_t frequency = std::min(documents_len, document_frequency);
    size_t diff = documents_len - frequency;
    Pre p;
    if (t.query_term_expanded == t.query_term) {
      p.expansion_boost = 1.0;
    } else {
      p.expansion_boost = std::log(
          1.0 + (1.0 / (1.0 + (double)t.query_term_expanded.size() - (double)t.query_term.size())));
    }
    p.idf = std::log(1.0 + ((double)diff + 0.5) / ((double)frequency + 0.5));
    return p;
  }

  // bm25.rs:60-93
  bool score(const Pre& pre, const DocumentPointer& dp, const DocumentDetails& dd, Aidx,
             const FieldData& fd, const TermData&, double* out) {
    double score = 0.0;
    for (size_t x = 0; x < dd.field_length.size(); ++x) {
      double tf = (double)dp.term_frequency[x];
      if (tf > 0.0) {
        double avg = fd.fields[x].avg;
        tf = ((bm25k1 + 1.0) * tf) /
             (bm25k1 * ((1.0 - bm25b) + bm25b * ((double)dd.field_length[x] / avg)) + tf);
        score += tf * pre.idf * fd.fields_boost[x] * pre.expansion_boost;
      }
    }
    if (score > 0.0) { *out = score; return true; }
    return false;
  }
  void finalize(std::vector<QueryResult>&) {}
};

// ----------------------------------------------------------------------------
// src/score/default/zero_to_one.rs
// ----------------------------------------------------------------------------
struct ZeroToOne {
  struct ScoreByTerm {           // zero_to_one.rs:27-34
    size_t query_term_index, all_query_terms_len, field_length, index_node_id, term_frequency;
    double score;
  };
  struct Pre {};
  FlatMap<std::vector<std::vector<ScoreByTerm>>> by_doc;   // zero_to_one.rs:24-26

  Pre before_each(const TermData&, size_t, size_t) { return Pre{}; }   // trait default → None

  // zero_to_one.rs:44-82
  bool score(const Pre&, const DocumentPointer& dp, const DocumentDetails& dd, Aidx node,
             const FieldData&, const TermData& t, double* out) {
    uint64_t key = dd.key;
    for (size_t x = 0; x < dd.field_length.size(); ++x) {
      size_t tf = dp.term_frequency[x];
      if (tf > 0) {
        double term_exp_len = (double)t.query_term_expanded.size();
        double term_len = (double)t.query_term.size();
        size_t field_length = dd.field_length[x];
        if (!by_doc.contains(key)) {
          std::vector<std::vector<ScoreByTerm>> v(dd.field_length.size());
          by_doc.insert(key, std::move(v));
        }
        (*by_doc.get(key))[x].push_back(ScoreByTerm{
            t.query_term_index, t.query_terms_len, field_length, (size_t)node.slot, tf,
            1.0 - std::fabs(term_exp_len - term_len) / term_exp_len});
      }
    }
    *out = 0.0;   // dummy Some(0.)
    return true;
  }

  // zero_to_one.rs:84-126
  void finalize(std::vector<QueryResult>& results) {
    for (QueryResult& r : results) {
      auto* fields = by_doc.get(r.key);
      for (auto& field_scores : *fields) {
        std::unordered_map<size_t, size_t> df_pool_by_id;
        std::unordered_map<size_t, char> consumed_index;
        // Rust sort_by is a STABLE sort
        std::stable_sort(field_scores.begin(), field_scores.end(),
                         [](const ScoreByTerm& a, const ScoreByTerm& b) { return a.score > b.score; });
        double score_by_pool = 0.0;
        for (const ScoreByTerm& s : field_scores) {
          if (consumed_index.count(s.query_term_index)) continue;
          auto it = df_pool_by_id.find(s.index_node_id);
          if (it != df_pool_by_id.end()) {
            if (it->second <= 0) continue;
            it->second -= 1;
          } else {
            df_pool_by_id.emplace(s.index_node_id, s.term_frequency - 1);
          }
          consumed_index.emplace(s.query_term_index, 1);
          double df = (double)s.term_frequency;
          score_by_pool += std::fmin(s.score / df, 1.0) * (double)s.term_frequency /
                           (double)std::max(s.field_length, s.all_query_terms_len);
        }
        r.score = std::fmax(score_by_pool, r.score);
      }
    }
    by_doc.clear();
  }
};

// Order-independent digests shared (by definition, not by code) with the GPU
// path: see include/probly_b200.h "Digests".
static inline uint32_t doc_mix(uint64_t doc) { return ((uint32_t)doc + 1u) * 0x9E3779B1u; }
static inline uint64_t doc_hash(uint64_t doc) { uint64_t a = doc_mix(doc); return a * a; }
static inline uint64_t score_hash(uint64_t doc, double score) {
  uint64_t b; std::memcpy(&b, &score, 8);
  uint32_t lo = (uint32_t)b, hi = (uint32_t)(b >> 32);
  uint64_t y = lo ^ (hi * 0x85EBCA77u) ^ doc_mix(doc);
  return y * y;
}

}  // namespace orc

// =============================================================================
// C ABI for the Python test harness (ctypes).  Tokens arrive pre-tokenized:
// a flat byte buffer + (n_tokens+1) offsets.
// =============================================================================
using namespace orc;

static std::string_view tok(const uint8_t* bytes, const uint64_t* off, uint64_t i) {
  return std::string_view((const char*)bytes + off[i], off[i + 1] - off[i]);
}

extern "C" {

void* orc_index_new(uint32_t fields_num) { return new Index(fields_num); }
void orc_index_free(void* h) { delete (Index*)h; }

// One document.  value_tok_count has one entry per field VALUE (how many tokens it
// tokenized to, empty tokens included); field_value_count[f] = number of values of field f.
void orc_add_document(void* h, uint64_t key, const uint8_t* tok_bytes, const uint64_t* tok_off,
                      const uint32_t* value_tok_count, const uint32_t* field_value_count) {
  Index* ix = (Index*)h;
  std::vector<std::vector<std::vector<std::string_view>>> fv(ix->fields.size());
  uint64_t t = 0, v = 0;
  for (size_t f = 0; f < ix->fields.size(); ++f) {
    for (uint32_t j = 0; j < field_value_count[f]; ++j, ++v) {
      std::vector<std::string_view> terms;
      for (uint32_t k = 0; k < value_tok_count[v]; ++k, ++t) terms.push_back(tok(tok_bytes, tok_off, t));
      fv[f].push_back(std::move(terms));
    }
  }
  ix->add_document(key, fv);
}

// Bulk: n_docs documents, every field has exactly ONE value; field_tok_count is
// [n_docs * F] token counts; tokens are concatenated in (doc, field) order.
void orc_add_documents(void* h, uint64_t n_docs, const uint64_t* keys, const uint8_t* tok_bytes,
                       const uint64_t* tok_off, const uint32_t* field_tok_count) {
  Index* ix = (Index*)h;
  const size_t F = ix->fields.size();
  uint64_t t = 0;
  std::vector<std::vector<std::vector<std::string_view>>> fv(F);
  for (uint64_t d = 0; d < n_docs; ++d) {
    for (size_t f = 0; f < F; ++f) {
      fv[f].assign(1, {});
      uint32_t n = field_tok_count[d * F + f];
      fv[f][0].reserve(n);
      for (uint32_t k = 0; k < n; ++k, ++t) fv[f][0].push_back(tok(tok_bytes, tok_off, t));
    }
    ix->add_document(keys[d], fv);
  }
}

void orc_remove_document(void* h, uint64_t key) { ((Index*)h)->remove_document(key); }
void orc_vacuum(void* h) { ((Index*)h)->vacuum(); }
uint64_t orc_docs_len(void* h) { return ((Index*)h)->docs.len(); }
uint64_t orc_arena_doc_len(void* h) { return ((Index*)h)->arena_doc.len(); }
uint64_t orc_arena_index_len(void* h) { return ((Index*)h)->arena_index.len(); }
uint64_t orc_score_calls(void*) { return tl_score_calls; }
void orc_field_stats(void* h, uint64_t* sum, double* avg) {
  Index* ix = (Index*)h;
  for (size_t i = 0; i < ix->fields.size(); ++i) { sum[i] = ix->fields[i].sum; avg[i] = ix->fields[i].avg; }
}

// count_nodes helper of the reference's tests (index.rs:464-480): nodes reachable from the root.
uint64_t orc_count_nodes(void* h) {
  Index* ix = (Index*)h;
  uint64_t n = 0;
  std::vector<Aidx> st{ix->root};
  while (!st.empty()) {
    Aidx a = st.back(); st.pop_back();
    ++n;
    const InvertedIndexNode* nd = ix->arena_index.get(a);
    if (nd->first_child.some()) st.push_back(nd->first_child);
    if (nd->next.some()) st.push_back(nd->next);
  }
  return n;
}

// Children chars of the node reached by `path`, in linked-list order (prepend order pins,
// index.rs:521-542).  Returns count; writes up to cap.
uint64_t orc_children_chars(void* h, const uint8_t* path, uint64_t path_len, uint32_t* out, uint64_t cap) {
  Index* ix = (Index*)h;
  Aidx n = ix->find_node(ix->root, std::string_view((const char*)path, path_len));
  if (!n.some()) return 0;
  uint64_t c = 0;
  Aidx it = ix->arena_index.get(n)->first_child;
  while (it.some()) {
    const InvertedIndexNode* nd = ix->arena_index.get(it);
    if (c < cap) out[c] = nd->ch;
    ++c;
    it = nd->next;
  }
  return c;
}

// Postings of a term in linked-list order: keys and the tf vector of each POINTER.
uint64_t orc_postings(void* h, const uint8_t* term, uint64_t term_len, uint64_t* keys, uint64_t* tf, uint64_t cap) {
  Index* ix = (Index*)h;
  Aidx n = ix->find_node(ix->root, std::string_view((const char*)term, term_len));
  if (!n.some()) return 0;
  uint64_t c = 0;
  const size_t F = ix->fields.size();
  Aidx it = ix->arena_index.get(n)->first_doc;
  while (it.some()) {
    const DocumentPointer* dp = ix->arena_doc.get(it);
    if (c < cap) {
      keys[c] = dp->details_key;
      for (size_t f = 0; f < F; ++f) tf[c * F + f] = dp->term_frequency[f];
    }
    ++c;
    it = dp->next;
  }
  return c;
}

// expand_term: writes the expansions joined by '\n' into out (cap bytes); returns the
// number of expansions, sets *needed to the bytes required.
uint64_t orc_expand_term(void* h, const uint8_t* term, uint64_t term_len, uint8_t* out, uint64_t cap,
                         uint64_t* needed) {
  Index* ix = (Index*)h;
  auto ex = ix->expand_term(std::string_view((const char*)term, term_len));
  uint64_t pos = 0;
  for (size_t i = 0; i < ex.size(); ++i) {
    for (char c : ex[i]) { if (pos < cap) out[pos] = (uint8_t)c; ++pos; }
    if (i + 1 < ex.size()) { if (pos < cap) out[pos] = '\n'; ++pos; }
  }
  *needed = pos;
  return ex.size();
}

static std::vector<QueryResult> run_query(Index* ix, const std::vector<std::string_view>& terms,
                                          int scorer, double k1, double b, const double* boosts) {
  if (scorer == 0) { BM25 c; c.bm25k1 = k1; c.bm25b = b; return ix->query(terms, c, boosts); }
  ZeroToOne z; return ix->query(terms, z, boosts);
}

// One query (pre-tokenized).  scorer: 0 = BM25, 1 = ZeroToOne.  Results come back in the
// reference's comparison order (score desc, key asc).  Returns the result count; writes up to cap.
uint64_t orc_query(void* h, const uint8_t* tok_bytes, const uint64_t* tok_off, uint64_t n_tokens,
                   int scorer, double k1, double b, const double* boosts, uint64_t* keys,
                   double* scores, uint64_t cap) {
  Index* ix = (Index*)h;
  std::vector<std::string_view> terms;
  for (uint64_t i = 0; i < n_tokens; ++i) terms.push_back(tok(tok_bytes, tok_off, i));
  auto r = run_query(ix, terms, scorer, k1, b, boosts);
  for (size_t i = 0; i < r.size() && i < cap; ++i) { keys[i] = r[i].key; scores[i] = r[i].score; }
  return r.size();
}

// Batch of queries, run entirely in C++ on n_threads host threads (queries statically
// partitioned; legal for the reference because `query` takes &self — query.rs:22 — and
// each thread owns its calculator).  Per query: result count, doc digest, score digest,
// top-k (key, score).  Returns elapsed wall seconds of the query loop only.
double orc_query_batch(void* h, uint64_t n_queries, const uint64_t* query_tok_off,
                       const uint8_t* tok_bytes, const uint64_t* tok_off, int scorer, double k1,
                       double b, const double* boosts, uint32_t top_k, uint32_t n_threads,
                       uint64_t* n_results, uint64_t* doc_digest, uint64_t* score_digest,
                       uint32_t* topk_n, uint64_t* topk_key, double* topk_score,
                       uint64_t* total_score_calls) {
  Index* ix = (Index*)h;
  if (n_threads == 0) n_threads = 1;
  std::vector<uint64_t> calls(n_threads, 0);
  auto work = [&](uint32_t tid) {
    uint64_t calls_before = tl_score_calls;
    uint64_t lo = n_queries * tid / n_threads, hi = n_queries * (tid + 1) / n_threads;
    for (uint64_t q = lo; q < hi; ++q) {
      std::vector<std::string_view> terms;
      for (uint64_t i = query_tok_off[q]; i < query_tok_off[q + 1]; ++i) terms.push_back(tok(tok_bytes, tok_off, i));
      auto r = run_query(ix, terms, scorer, k1, b, boosts);
      uint64_t dd = 0, sd = 0;
      for (auto& e : r) { dd += doc_hash(e.key); sd += score_hash(e.key, e.score); }
      if (n_results) n_results[q] = r.size();
      if (doc_digest) doc_digest[q] = dd;
      if (score_digest) score_digest[q] = sd;
      uint32_t n = (uint32_t)std::min<uint64_t>(top_k, r.size());
      if (topk_n) topk_n[q] = n;
      for (uint32_t i = 0; i < n; ++i) {
        if (topk_key) topk_key[q * top_k + i] = r[i].key;
        if (topk_score) topk_score[q * top_k + i] = r[i].score;
      }
    }
    calls[tid] = tl_score_calls - calls_before;
  };
  auto t0 = std::chrono::steady_clock::now();
  if (n_threads == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < n_threads; ++t) th.emplace_back(work, t);
    for (auto& t : th) t.join();
  }
  auto t1 = std::chrono::steady_clock::now();
  if (total_score_calls) { uint64_t t = 0; for (uint64_t c : calls) t += c; *total_score_calls = t; }
  return std::chrono::duration<double>(t1 - t0).count();
}

uint64_t orc_doc_hash(uint64_t doc) { return doc_hash(doc); }
uint64_t orc_score_hash(uint64_t doc, double s) { return score_hash(doc, s); }

}  // extern "C"
