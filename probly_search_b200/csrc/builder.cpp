// Host-side index maintenance + flattener.
//
// Replaces — as SEMANTICS, with a different data layout — the reference's L1
// (`Index<T>`: src/index.rs:19-33) for the mutation calls the query path depends on:
//   add_document    src/index.rs:77-158      remove_document  src/index.rs:161-191
//   vacuum          src/index.rs:194-241
// The reference keeps arenas of linked trie nodes and one linked posting node per term
// OCCURRENCE, tuned for cheap mutation.  Here the builder keeps
//   * a term dictionary (bytes -> term id) so repeated terms skip the trie walk,
//   * trie nodes in creation order with a (parent, char) hash for descent,
//   * an append-only log of (term, doc, tf[F]) tuples, one per (doc, DISTINCT term),
//   * a doc table (key, field_length[F], state),
// and `flatten()` turns that into the immutable image the GPU reads: a CSR trie renumbered in
// DFS pre-order (children most-recently-created first, the order src/index.rs:409-419
// produces by prepending) and term-major SoA posting columns (pb_index_image in
// include/probly_b200.h).
//
// Behaviour that defines the data the query path reads and is reproduced exactly
// (SURVEY.md §3.4): rule 1/2 (multiplicity = sum of tf), rule 5 (expansion order), rule 7
// (byte lengths), rule 8 (field_length = token count of the LAST value, sum over all values,
// avg updated per value with docs.len()+1, removal arithmetic incl. NaN on the last doc).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <string_view>
#include <vector>

#include "common.hpp"

namespace pb {

static inline uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}
static inline uint64_t hash_bytes(const uint8_t* p, size_t n) {
  uint64_t h = 0x9E3779B97F4A7C15ULL ^ (n * 0xD6E8FEB86659FD93ULL);
  while (n >= 8) { uint64_t w; std::memcpy(&w, p, 8); h = mix64(h ^ w) ; p += 8; n -= 8; }
  uint64_t w = 0;
  std::memcpy(&w, p, n);
  return mix64(h ^ w ^ (uint64_t(n) << 56));
}

// u64 -> u32 open-addressing map (keys: doc keys, (parent,char) pairs).  EMPTY value = absent.
class U64Map {
  std::vector<uint64_t> k_;
  std::vector<uint32_t> v_;
  size_t len_ = 0, mask_ = 0;
  static constexpr uint32_t kEmpty = 0xFFFFFFFFu, kTomb = 0xFFFFFFFEu;
  void rehash(size_t ncap) {
    std::vector<uint64_t> ok; std::vector<uint32_t> ov;
    ok.swap(k_); ov.swap(v_);
    k_.assign(ncap, 0); v_.assign(ncap, kEmpty);
    mask_ = ncap - 1; len_ = 0; used_ = 0;
    for (size_t i = 0; i < ov.size(); ++i) if (ov[i] < kTomb) put(ok[i], ov[i]);
  }
  size_t used_ = 0;   // live + tombstones

 public:
  size_t size() const { return len_; }
  uint32_t get(uint64_t key) const {
    if (k_.empty()) return kEmpty;
    size_t i = mix64(key) & mask_;
    while (v_[i] != kEmpty) {
      if (v_[i] != kTomb && k_[i] == key) return v_[i];
      i = (i + 1) & mask_;
    }
    return kEmpty;
  }
  void put(uint64_t key, uint32_t val) {
    if ((used_ + 1) * 10 > k_.size() * 6) rehash(k_.empty() ? 16 : (len_ * 4 > k_.size() ? k_.size() * 2 : k_.size()));
    size_t i = mix64(key) & mask_;
    size_t tomb = SIZE_MAX;
    while (v_[i] != kEmpty) {
      if (v_[i] == kTomb) { if (tomb == SIZE_MAX) tomb = i; }
      else if (k_[i] == key) { v_[i] = val; return; }
      i = (i + 1) & mask_;
    }
    if (tomb != SIZE_MAX) i = tomb; else ++used_;
    k_[i] = key; v_[i] = val; ++len_;
  }
  bool erase(uint64_t key) {
    if (k_.empty()) return false;
    size_t i = mix64(key) & mask_;
    while (v_[i] != kEmpty) {
      if (v_[i] != kTomb && k_[i] == key) { v_[i] = kTomb; --len_; return true; }
      i = (i + 1) & mask_;
    }
    return false;
  }
  static constexpr uint32_t npos = kEmpty;
};

// bytes -> term id dictionary; strings live in one arena.
class TermDict {
  std::vector<uint32_t> slot_;   // term id or EMPTY
  size_t mask_ = 0;
  static constexpr uint32_t kEmpty = 0xFFFFFFFFu;

 public:
  std::vector<uint8_t> bytes;
  std::vector<uint64_t> off{0};
  size_t size() const { return off.size() - 1; }
  std::string_view str(uint32_t t) const { return {(const char*)bytes.data() + off[t], size_t(off[t + 1] - off[t])}; }
  void grow() {
    size_t ncap = slot_.empty() ? 1024 : slot_.size() * 2;
    slot_.assign(ncap, kEmpty);
    mask_ = ncap - 1;
    for (uint32_t t = 0; t < size(); ++t) {
      size_t i = hash_bytes(bytes.data() + off[t], off[t + 1] - off[t]) & mask_;
      while (slot_[i] != kEmpty) i = (i + 1) & mask_;
      slot_[i] = t;
    }
  }
  // returns (term id, inserted?)
  std::pair<uint32_t, bool> intern(const uint8_t* p, size_t n) {
    if ((size() + 1) * 2 > slot_.size()) grow();
    size_t i = hash_bytes(p, n) & mask_;
    while (slot_[i] != kEmpty) {
      uint32_t t = slot_[i];
      if (off[t + 1] - off[t] == n && std::memcmp(bytes.data() + off[t], p, n) == 0) return {t, false};
      i = (i + 1) & mask_;
    }
    uint32_t t = (uint32_t)size();
    bytes.insert(bytes.end(), p, p + n);
    off.push_back(bytes.size());
    slot_[i] = t;
    return {t, true};
  }
};

static inline bool utf8_next(const uint8_t*& p, const uint8_t* end, uint32_t* cp) {
  uint8_t c = *p++;
  if (c < 0x80) { *cp = c; return true; }
  int extra;
  uint32_t v;
  if ((c & 0xE0) == 0xC0) { extra = 1; v = c & 0x1F; }
  else if ((c & 0xF0) == 0xE0) { extra = 2; v = c & 0x0F; }
  else if ((c & 0xF8) == 0xF0) { extra = 3; v = c & 0x07; }
  else return false;
  if (end - p < extra) return false;
  for (int i = 0; i < extra; ++i) {
    uint8_t d = *p++;
    if ((d & 0xC0) != 0x80) return false;
    v = (v << 6) | (d & 0x3F);
  }
  *cp = v;
  return true;
}

bool utf8_valid(const uint8_t* p, size_t n) {
  const uint8_t* e = p + n;
  uint32_t cp;
  while (p < e) if (!utf8_next(p, e, &cp)) return false;
  return true;
}

struct Node {
  uint32_t ch;
  uint32_t parent;
  uint32_t term;     // term id attached to this node or NONE
  bool alive;
};
static constexpr uint32_t NONE = 0xFFFFFFFFu;

enum DocState : uint8_t { LIVE = 0, REMOVED_PENDING = 1, GONE = 2 };

struct Tuple {            // one per (doc, distinct term)
  uint32_t term;
  uint32_t doc;
  uint32_t tf[PB_MAX_FIELDS];
};

struct Builder {
  uint32_t F;
  // docs
  std::vector<uint64_t> doc_key;
  std::vector<uint32_t> doc_fl;          // [n * F]
  std::vector<uint8_t> doc_state;
  std::vector<uint64_t> doc_log_begin;   // [n + 1] range of the doc's tuples in `log`
  U64Map key2ord;                        // live docs only (removal deletes from `docs`, index.rs:188-190)
  U64Map removed_keys;                   // removed-but-not-vacuumed keys (index.rs:32)
  uint64_t n_live = 0, n_removed_pending = 0;
  uint64_t field_sum[PB_MAX_FIELDS] = {0, 0, 0, 0};
  double field_avg[PB_MAX_FIELDS] = {0, 0, 0, 0};
  // terms + trie
  TermDict dict;
  std::vector<uint32_t> term_node;       // term id -> node (NONE when its node was pruned)
  std::vector<uint64_t> term_rows;       // tuples currently in the log for this term
  std::vector<uint32_t> term_stamp;      // dedupe inside one document
  std::vector<uint64_t> term_logpos;
  std::vector<Node> nodes;               // creation order; nodes[0] = root ('\0', index.rs:54)
  U64Map child;                          // (parent << 21 | char) -> node
  uint64_t n_alive_nodes = 1;
  std::vector<Tuple> log;
  uint64_t n_pointers = 0;               // sum of multiplicities of tuples in the log

  // flatten outputs (owned here, handed out as raw pointers)
  bool flat_valid = false;
  bool flat_posts = false;               // the flattened image carries the host posting columns (post_blocks)
  uint64_t flat_from = 0;                // first doc ordinal whose rows the flattened image holds (0 = all: the full image)
  std::vector<uint32_t> f_term_id;       // DFS term ordinal of the last flatten -> builder term id (stable across flattens)
  std::vector<uint32_t> f_node_edge_begin, f_node_term_lo, f_node_term_hi, f_node_parent, f_node_char;
  std::vector<uint32_t> f_edge_char, f_edge_child;
  std::vector<uint64_t> f_term_row_begin;
  std::vector<uint32_t> f_term_byte_len, f_term_node;
  std::vector<uint32_t> f_post_blocks;   // tile-blocked columns, see pb_index_image
  std::vector<uint32_t> f_removed;
  pb_index_image image{};

  explicit Builder(uint32_t f) : F(f) {
    nodes.push_back(Node{0, NONE, NONE, true});
    doc_log_begin.push_back(0);
  }

  static uint64_t ckey(uint32_t parent, uint32_t ch) { return (uint64_t(parent) << 21) | ch; }

  // Trie descent with node creation (index.rs:119-147 + 437-452): returns the node of `term`.
  uint32_t descend_create(const uint8_t* p, size_t n) {
    const uint8_t* e = p + n;
    uint32_t cur = 0, cp = 0;
    while (p < e) {
      utf8_next(p, e, &cp);
      uint32_t nx = child.get(ckey(cur, cp));
      if (nx == U64Map::npos) {
        nx = (uint32_t)nodes.size();
        nodes.push_back(Node{cp, cur, NONE, true});   // creation order == id order
        child.put(ckey(cur, cp), nx);
        ++n_alive_nodes;
      }
      cur = nx;
    }
    return cur;
  }

  int add_document(uint64_t key, const uint8_t* tb, const uint64_t* to,
                   const uint32_t* value_tok_count, const uint32_t* field_value_count) {
    if (key2ord.get(key) != U64Map::npos) {
      set_error("add_document: key %llu is already in the index (re-adding a live key is undefined in the reference, SURVEY §3.4 rule 13)", (unsigned long long)key);
      return PB_ERR_DUPLICATE_KEY;
    }
    if (removed_keys.get(key) != U64Map::npos) {
      set_error("add_document: key %llu was removed but not vacuumed; vacuum first", (unsigned long long)key);
      return PB_ERR_DUPLICATE_KEY;
    }
    if (doc_key.size() >= 0xFFFFFFF0ull) { set_error("too many documents"); return PB_ERR_UNSUPPORTED; }
    // validate first so a failure leaves the builder untouched
    {
      uint64_t t = 0, v = 0;
      for (uint32_t f = 0; f < F; ++f)
        for (uint32_t j = 0; j < field_value_count[f]; ++j, ++v)
          for (uint32_t k = 0; k < value_tok_count[v]; ++k, ++t) {
            if (to[t + 1] < to[t]) { set_error("add_document: token offsets not monotone"); return PB_ERR_INVALID; }
            if (!utf8_valid(tb + to[t], to[t + 1] - to[t])) { set_error("add_document: token %llu is not valid UTF-8", (unsigned long long)t); return PB_ERR_INVALID; }
          }
    }
    flat_valid = false;
    const uint32_t ord = (uint32_t)doc_key.size();
    uint32_t fl[PB_MAX_FIELDS] = {0, 0, 0, 0};
    uint64_t t = 0, v = 0;
    for (uint32_t f = 0; f < F; ++f) {
      for (uint32_t j = 0; j < field_value_count[f]; ++j, ++v) {
        uint32_t filtered = 0;
        for (uint32_t k = 0; k < value_tok_count[v]; ++k, ++t) {
          const size_t n = to[t + 1] - to[t];
          if (n == 0) continue;                                 // index.rs:101
          ++filtered;
          auto [tid, fresh] = dict.intern(tb + to[t], n);
          if (fresh) {
            term_node.push_back(NONE); term_rows.push_back(0);
            term_stamp.push_back(0); term_logpos.push_back(0);
          }
          if (term_node[tid] == NONE || !nodes[term_node[tid]].alive) {
            uint32_t nd = descend_create(tb + to[t], n);
            term_node[tid] = nd;
            nodes[nd].term = tid;
          }
          if (term_stamp[tid] != ord + 1) {                     // first occurrence in this doc
            term_stamp[tid] = ord + 1;
            term_logpos[tid] = log.size();
            Tuple tp{tid, ord, {0, 0, 0, 0}};
            log.push_back(tp);
            ++term_rows[tid];
          }
          ++log[term_logpos[tid]].tf[f];
          ++n_pointers;
        }
        field_sum[f] += filtered;                                       // index.rs:112
        field_avg[f] = (double)field_sum[f] / ((double)n_live + 1.0);   // index.rs:113
        fl[f] = filtered;                                               // index.rs:114 (last value wins)
      }
    }
    doc_key.push_back(key);
    for (uint32_t f = 0; f < F; ++f) doc_fl.push_back(fl[f]);
    doc_state.push_back(LIVE);
    doc_log_begin.push_back(log.size());
    key2ord.put(key, ord);                                              // index.rs:118
    ++n_live;
    return PB_OK;
  }

  // index.rs:161-191
  int remove_document(uint64_t key) {
    uint32_t ord = key2ord.get(key);
    if (ord == U64Map::npos) return PB_OK;            // unknown key: the reference does nothing either
    removed_keys.put(key, ord);
    double new_len = (double)(n_live - 1);
    for (uint32_t f = 0; f < F; ++f) {
      uint32_t fl = doc_fl[size_t(ord) * F + f];
      if (fl > 0) {
        field_sum[f] -= fl;
        field_avg[f] = (double)field_sum[f] / new_len;   // NaN when the last doc goes (index.rs:643)
      }
    }
    doc_state[ord] = REMOVED_PENDING;
    key2ord.erase(key);
    --n_live;
    ++n_removed_pending;
    // Removal is lazy in the reference too (the postings stay until vacuum): the flattened structure is
    // untouched, only the image's live state follows — O(1), no re-flatten of the posting columns.
    if (flat_valid) {
      f_removed[ord >> 5] |= 1u << (ord & 31);
      image.n_removed = n_removed_pending;
      image.n_live_docs = n_live;
      for (uint32_t f = 0; f < F; ++f) image.field_avg[f] = field_avg[f];
    }
    return PB_OK;
  }

  // index.rs:194-241: drop the removed docs' postings and prune subtrees that own no posting.
  int vacuum() {
    flat_valid = false;
    if (n_removed_pending) {
      std::vector<Tuple> keep;
      keep.reserve(log.size());
      std::vector<uint64_t> nb(doc_log_begin.size(), 0);
      for (size_t d = 0; d + 1 < doc_log_begin.size(); ++d) {
        nb[d] = keep.size();
        if (doc_state[d] == REMOVED_PENDING) {
          for (uint64_t i = doc_log_begin[d]; i < doc_log_begin[d + 1]; ++i) {
            --term_rows[log[i].term];
            for (uint32_t f = 0; f < F; ++f) n_pointers -= log[i].tf[f];
          }
          doc_state[d] = GONE;
        } else {
          for (uint64_t i = doc_log_begin[d]; i < doc_log_begin[d + 1]; ++i) keep.push_back(log[i]);
        }
      }
      nb.back() = keep.size();
      log.swap(keep);
      doc_log_begin.swap(nb);
      n_removed_pending = 0;
    }
    removed_keys = U64Map();
    // prune: a node survives iff its subtree owns a posting (vacuum_node's return value)
    std::vector<uint8_t> has(nodes.size(), 0);
    for (size_t i = nodes.size(); i-- > 0;) {          // children have larger ids than parents
      Node& nd = nodes[i];
      if (!nd.alive) continue;
      if (nd.term != NONE && term_rows[nd.term] > 0) has[i] = 1;
      if (has[i] && nd.parent != NONE) has[nd.parent] = 1;
    }
    for (size_t i = 1; i < nodes.size(); ++i) {
      Node& nd = nodes[i];
      if (nd.alive && !has[i]) {
        nd.alive = false;
        child.erase(ckey(nd.parent, nd.ch));
        if (nd.term != NONE) { term_node[nd.term] = NONE; nd.term = NONE; }
        --n_alive_nodes;
      }
    }
    return PB_OK;
  }

  void info(pb_builder_info* o) const {
    std::memset(o, 0, sizeof(*o));
    o->num_fields = F;
    o->n_live_docs = n_live;
    o->n_doc_ordinals = doc_key.size();
    o->n_removed_pending = n_removed_pending;
    uint64_t nt = 0;
    for (uint64_t r : term_rows) nt += (r > 0);
    o->n_terms = nt;
    o->n_nodes = n_alive_nodes;
    o->n_rows = log.size();
    o->n_pointers = n_pointers;
    for (uint32_t f = 0; f < F; ++f) { o->field_sum[f] = field_sum[f]; o->field_avg[f] = field_avg[f]; }
  }

  // from_doc = 0: the whole index.  from_doc > 0: a DELTA segment — the current trie (so that its expansion order is
  // the global one) with the posting rows of the docs whose ordinal is >= from_doc only.
  // with_posts = false: everything but the posting columns (post_blocks = NULL, max_tf / max_fl = 0) — the device
  // builds the columns itself from the append log (pb_index_create_from_builder).
  int flatten(pb_index_image* out, uint64_t from_doc = 0, bool with_posts = true) {
    if (from_doc > doc_key.size()) { set_error("flatten: first doc ordinal %llu beyond the index", (unsigned long long)from_doc); return PB_ERR_INVALID; }
    if (!flat_valid || flat_from != from_doc || (with_posts && !flat_posts)) {
      flat_valid = false;
      int rc = do_flatten(from_doc, with_posts);
      if (rc != PB_OK) return rc;
      flat_valid = true;
      flat_posts = with_posts;
      flat_from = from_doc;
    }
    *out = image;
    if (!with_posts) { out->post_blocks = nullptr; for (uint32_t f = 0; f < PB_MAX_FIELDS; ++f) { out->max_tf[f] = 0; out->max_fl[f] = 0; } }
    return PB_OK;
  }

  std::vector<uint32_t> f_ord_of;        // builder term id -> DFS term ordinal of the last flatten (NONE: not in it)

  int do_flatten(uint64_t from_doc, bool with_posts) {
    const size_t NN = nodes.size();
    // rows per term inside the flattened doc range
    const uint64_t log_from = doc_log_begin[from_doc];
    std::vector<uint64_t> rows_local;
    if (from_doc) {
      rows_local.assign(dict.size(), 0);
      for (uint64_t i = log_from; i < log.size(); ++i) ++rows_local[log[i].term];
    }
    const std::vector<uint64_t>& rows_of = from_doc ? rows_local : term_rows;
    // children per parent, most recently created first (ids descend)
    std::vector<uint32_t> cnt(NN + 1, 0);
    for (size_t i = 1; i < NN; ++i) if (nodes[i].alive) ++cnt[nodes[i].parent + 1];
    for (size_t i = 0; i < NN; ++i) cnt[i + 1] += cnt[i];
    std::vector<uint32_t> kids(cnt[NN]);
    {
      std::vector<uint32_t> fill(cnt.begin(), cnt.end() - 1);
      for (size_t i = NN; i-- > 1;) if (nodes[i].alive) kids[fill[nodes[i].parent]++] = (uint32_t)i;
    }
    // DFS pre-order renumbering
    const size_t NA = n_alive_nodes;
    std::vector<uint32_t> new_id(NN, NONE), order;
    order.reserve(NA);
    f_node_term_lo.assign(NA, 0); f_node_term_hi.assign(NA, 0);
    f_node_parent.assign(NA, NONE); f_node_char.assign(NA, 0);
    f_term_node.clear(); f_term_byte_len.clear();
    std::vector<uint32_t> term_old;            // DFS term ordinal -> builder term id
    struct Frame { uint32_t node; uint32_t next_kid; };
    std::vector<Frame> st;
    auto enter = [&](uint32_t old) {
      uint32_t id = (uint32_t)order.size();
      new_id[old] = id;
      order.push_back(old);
      f_node_term_lo[id] = (uint32_t)term_old.size();
      const Node& nd = nodes[old];
      f_node_char[id] = nd.ch;
      f_node_parent[id] = nd.parent == NONE ? NONE : new_id[nd.parent];
      if (nd.term != NONE && rows_of[nd.term] > 0) {         // first_doc.is_some() (query.rs:136)
        f_term_node.push_back(id);
        f_term_byte_len.push_back((uint32_t)dict.str(nd.term).size());
        term_old.push_back(nd.term);
      }
      st.push_back(Frame{old, cnt[old]});
    };
    enter(0);
    while (!st.empty()) {
      Frame& fr = st.back();
      if (fr.next_kid < cnt[fr.node + 1]) {
        uint32_t k = kids[fr.next_kid++];
        enter(k);
      } else {
        f_node_term_hi[new_id[fr.node]] = (uint32_t)term_old.size();
        st.pop_back();
      }
    }
    const size_t NT = term_old.size();
    // CSR edges sorted by char
    f_node_edge_begin.assign(NA + 1, 0);
    for (size_t id = 0; id < NA; ++id) {
      uint32_t old = order[id];
      f_node_edge_begin[id + 1] = f_node_edge_begin[id] + (cnt[old + 1] - cnt[old]);
    }
    const size_t NE = f_node_edge_begin[NA];
    f_edge_char.assign(NE, 0); f_edge_child.assign(NE, 0);
    {
      std::vector<std::pair<uint32_t, uint32_t>> tmp;
      for (size_t id = 0; id < NA; ++id) {
        uint32_t old = order[id];
        tmp.clear();
        for (uint32_t j = cnt[old]; j < cnt[old + 1]; ++j) tmp.emplace_back(nodes[kids[j]].ch, new_id[kids[j]]);
        std::sort(tmp.begin(), tmp.end());
        uint32_t b = f_node_edge_begin[id];
        for (size_t j = 0; j < tmp.size(); ++j) { f_edge_char[b + j] = tmp[j].first; f_edge_child[b + j] = tmp[j].second; }
      }
    }
    // posting columns: counting sort of the log by DFS term ordinal (stable -> docs ascend)
    std::vector<uint32_t>& ord_of = f_ord_of;
    ord_of.assign(dict.size(), NONE);
    for (size_t t = 0; t < NT; ++t) ord_of[term_old[t]] = (uint32_t)t;
    f_term_row_begin.assign(NT + 1, 0);
    for (size_t t = 0; t < NT; ++t) f_term_row_begin[t + 1] = f_term_row_begin[t] + rows_of[term_old[t]];
    const uint64_t NR = f_term_row_begin[NT];
    if (NR != log.size() - log_from) { set_error("flatten: internal row count mismatch"); return PB_ERR_INVALID; }
    f_term_id = term_old;
    const uint64_t NRP = ((NR + 127) / 128 + 1) * 128;     // pad: whole 128-row tiles + one spare tile
    const uint32_t NCOL = 1 + 2 * F;
    uint32_t max_tf[PB_MAX_FIELDS] = {0, 0, 0, 0}, max_fl[PB_MAX_FIELDS] = {0, 0, 0, 0};
    if (!with_posts) {
      std::vector<uint32_t>().swap(f_post_blocks);      // the device builds the columns from the log
    } else {
      f_post_blocks.assign(NRP * NCOL, 0);
      std::vector<uint64_t> fill(f_term_row_begin.begin(), f_term_row_begin.end() - 1);
      for (uint64_t li = log_from; li < log.size(); ++li) {
        const Tuple& tp = log[li];
        uint64_t r = fill[ord_of[tp.term]]++;
        uint32_t* blk = f_post_blocks.data() + (r / 128) * (uint64_t)NCOL * 128 + (r % 128);
        blk[0] = tp.doc;
        for (uint32_t f = 0; f < F; ++f) {
          uint32_t fl = doc_fl[size_t(tp.doc) * F + f];
          blk[(1 + f) * 128] = tp.tf[f];
          blk[(1 + F + f) * 128] = fl;
          max_tf[f] = std::max(max_tf[f], tp.tf[f]);
          max_fl[f] = std::max(max_fl[f], fl);
        }
      }
    }
    // live state
    const size_t ND = doc_key.size();
    f_removed.assign((ND + 31) / 32 + 1, 0);
    uint64_t nrem = 0;
    for (size_t d = 0; d < ND; ++d)      // ordinals vacuum left behind (GONE) own no posting row: only pending removals need the mask
      if (doc_state[d] == REMOVED_PENDING) { f_removed[d >> 5] |= 1u << (d & 31); ++nrem; }

    pb_index_image& im = image;
    std::memset(&im, 0, sizeof(im));
    im.version = 1;
    im.num_fields = F;
    im.n_nodes = NA; im.n_edges = NE; im.n_terms = NT; im.n_rows = NR; im.n_rows_padded = NRP; im.n_docs = ND;
    uint32_t mtb = 0;
    for (uint32_t x : f_term_byte_len) mtb = std::max(mtb, x);
    im.max_term_bytes = mtb;
    for (uint32_t f = 0; f < F; ++f) { im.max_tf[f] = max_tf[f]; im.max_fl[f] = max_fl[f]; }
    im.node_edge_begin = f_node_edge_begin.data();
    im.node_term_lo = f_node_term_lo.data(); im.node_term_hi = f_node_term_hi.data();
    im.node_parent = f_node_parent.data(); im.node_char = f_node_char.data();
    im.edge_char = f_edge_char.data(); im.edge_child = f_edge_child.data();
    im.term_row_begin = f_term_row_begin.data();
    im.term_byte_len = f_term_byte_len.data(); im.term_node = f_term_node.data();
    im.post_blocks = f_post_blocks.data();
    im.doc_key = doc_key.data();
    im.removed_bitmap = f_removed.data();
    im.n_removed = nrem;
    im.n_live_docs = n_live;
    for (uint32_t f = 0; f < F; ++f) im.field_avg[f] = field_avg[f];
    return PB_OK;
  }
};

}  // namespace pb

struct pb_builder { pb::Builder impl; explicit pb_builder(uint32_t f) : impl(f) {} };

namespace pb {
static_assert(sizeof(Tuple) == sizeof(LogTuple) && alignof(Tuple) == alignof(LogTuple), "the log tuple is shared with the engine");
int builder_flatten_structure(pb_builder* b, uint64_t from_doc, pb_index_image* im, BuilderLogView* view) {
  Builder& B = b->impl;
  int rc = B.flatten(im, from_doc, false);
  if (rc != PB_OK) return rc;
  const uint64_t log_from = B.doc_log_begin[from_doc];
  view->F = B.F;
  view->tuples = reinterpret_cast<const LogTuple*>(B.log.data()) + log_from;
  view->n_tuples = B.log.size() - log_from;
  view->doc_fl = B.doc_fl.data(); view->n_docs = B.doc_key.size();
  view->ord_of = B.f_ord_of.data(); view->n_ord = B.f_ord_of.size();
  return PB_OK;
}
}  // namespace pb

extern "C" {

int pb_builder_create(uint32_t num_fields, pb_builder** out) {
  if (!out) return PB_ERR_INVALID;
  if (num_fields == 0 || num_fields > PB_MAX_FIELDS) {
    pb::set_error("pb_builder_create: num_fields must be 1..%u", PB_MAX_FIELDS);
    return PB_ERR_UNSUPPORTED;
  }
  PB_TRY({ *out = new pb_builder(num_fields); return PB_OK; });
}

void pb_builder_destroy(pb_builder* b) { delete b; }

int pb_builder_add_document(pb_builder* b, uint64_t key, const pb_doc_tokens* d) {
  if (!b || !d || !d->tok_off || !d->field_value_count) { pb::set_error("pb_builder_add_document: null argument"); return PB_ERR_INVALID; }
  PB_TRY({ return b->impl.add_document(key, d->tok_bytes, d->tok_off, d->value_tok_count, d->field_value_count); });
}

int pb_builder_add_documents(pb_builder* b, uint64_t n_docs, const uint64_t* keys, const uint8_t* tok_bytes,
                             const uint64_t* tok_off, const uint32_t* field_tok_count) {
  if (!b || (n_docs && (!keys || !tok_off || !field_tok_count))) { pb::set_error("pb_builder_add_documents: null argument"); return PB_ERR_INVALID; }
  PB_TRY({
    const uint32_t F = b->impl.F;
    uint32_t ones[PB_MAX_FIELDS] = {1, 1, 1, 1};
    uint64_t t = 0;
    for (uint64_t d = 0; d < n_docs; ++d) {
      const uint32_t* vc = field_tok_count + d * F;
      int rc = b->impl.add_document(keys[d], tok_bytes, tok_off + t, vc, ones);
      if (rc != PB_OK) return rc;
      for (uint32_t f = 0; f < F; ++f) t += vc[f];
    }
    return PB_OK;
  });
}

int pb_builder_remove_document(pb_builder* b, uint64_t key) {
  if (!b) return PB_ERR_INVALID;
  PB_TRY({ return b->impl.remove_document(key); });
}

int pb_builder_vacuum(pb_builder* b) {
  if (!b) return PB_ERR_INVALID;
  PB_TRY({ return b->impl.vacuum(); });
}

int pb_builder_get_info(const pb_builder* b, pb_builder_info* out) {
  if (!b || !out) return PB_ERR_INVALID;
  PB_TRY({ b->impl.info(out); return PB_OK; });
}

int pb_builder_flatten_from(pb_builder* b, uint64_t from_doc_ordinal, pb_index_image* out) {
  if (!b || !out) { pb::set_error("pb_builder_flatten_from: null argument"); return PB_ERR_INVALID; }
  PB_TRY({ return b->impl.flatten(out, from_doc_ordinal); });
}

int pb_builder_flatten_term_ids(const pb_builder* b, uint32_t* out, uint64_t cap, uint64_t* n_terms) {
  if (!b || !n_terms) { pb::set_error("pb_builder_flatten_term_ids: null argument"); return PB_ERR_INVALID; }
  if (!b->impl.flat_valid) { pb::set_error("pb_builder_flatten_term_ids: no flattened image (call pb_builder_flatten first)"); return PB_ERR_INVALID; }
  *n_terms = b->impl.f_term_id.size();
  if (cap < b->impl.f_term_id.size()) { pb::set_error("pb_builder_flatten_term_ids: need %llu entries", (unsigned long long)b->impl.f_term_id.size()); return PB_ERR_CAPACITY; }
  if (out && !b->impl.f_term_id.empty()) std::memcpy(out, b->impl.f_term_id.data(), b->impl.f_term_id.size() * sizeof(uint32_t));
  return PB_OK;
}

int pb_builder_flatten(pb_builder* b, pb_index_image* out) {
  if (!b || !out) return PB_ERR_INVALID;
  PB_TRY({ return b->impl.flatten(out); });
}

}  // extern "C"
