// On-disk / wire format of the flattened index image (the reference has no
// serialisation at all — no serde, the index lives only in RAM).  One file = the pb_index_image a
// builder flattens: header, section table, 64-byte aligned sections, FNV-1a checksum of the payload.
// A loaded file yields a pb_index_image whose pointers point into the file's buffer, ready for
// pb_index_create: a process can serve queries without ever holding the mutable host index.
//
//   offset 0   magic "PBIMG2\0\0"            8 bytes
//          8   header_bytes (u32)  n_sections (u32)
//         16   scalar block: the non-pointer fields of pb_index_image, little endian, in order
//              (version, num_fields, n_nodes, n_edges, n_terms, n_rows, n_rows_padded, n_docs,
//               max_term_bytes, max_tf[4], max_fl[4], n_removed, n_live_docs, field_avg[4])
//          ..  section table: n_sections x {offset u64, bytes u64}
//          ..  checksum (u64, FNV-1a 64 over the scalar block, then all section bytes in table order)
//          ..  sections, each starting on a 64-byte boundary
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

#include "../../include/probly_b200.h"
#include "common.hpp"

namespace {

constexpr char MAGIC[8] = {'P', 'B', 'I', 'M', 'G', '2', 0, 0};
constexpr uint32_t N_SECTIONS = 13;

struct Scalars {
  uint32_t version, num_fields;
  uint64_t n_nodes, n_edges, n_terms, n_rows, n_rows_padded, n_docs;
  uint32_t max_term_bytes, pad0;
  uint32_t max_tf[PB_MAX_FIELDS], max_fl[PB_MAX_FIELDS];
  uint64_t n_removed, n_live_docs;
  double field_avg[PB_MAX_FIELDS];
};

struct Section { const void* p; uint64_t bytes; };

void sections_of(const pb_index_image& im, Section (&s)[N_SECTIONS]) {
  const uint64_t F = im.num_fields;
  s[0] = {im.node_edge_begin, (im.n_nodes + 1) * 4};
  s[1] = {im.node_term_lo, im.n_nodes * 4};
  s[2] = {im.node_term_hi, im.n_nodes * 4};
  s[3] = {im.node_parent, im.n_nodes * 4};
  s[4] = {im.node_char, im.n_nodes * 4};
  s[5] = {im.edge_char, im.n_edges * 4};
  s[6] = {im.edge_child, im.n_edges * 4};
  s[7] = {im.term_row_begin, (im.n_terms + 1) * 8};
  s[8] = {im.term_byte_len, im.n_terms * 4};
  s[9] = {im.term_node, im.n_terms * 4};
  s[10] = {im.post_blocks, im.n_rows_padded * (1 + 2 * F) * 4};
  s[11] = {im.doc_key, im.n_docs * 8};
  s[12] = {im.removed_bitmap, ((im.n_docs + 31) / 32 + 1) * 4};
}

uint64_t fnv1a(uint64_t h, const void* p, uint64_t n) {
  const uint8_t* b = static_cast<const uint8_t*>(p);
  // 8 bytes per step keeps a 0.5 GB image at a fraction of a second; still byte-order defined
  uint64_t i = 0;
  for (; i + 8 <= n; i += 8) { uint64_t w; std::memcpy(&w, b + i, 8); h = (h ^ w) * 0x100000001B3ull; }
  for (; i < n; ++i) h = (h ^ b[i]) * 0x100000001B3ull;
  return h;
}

uint64_t align64(uint64_t x) { return (x + 63) & ~uint64_t(63); }

// Range checks before any size is multiplied: a corrupt header must not overflow a size product.
bool scalars_sane(const pb_index_image& im) {
  const uint64_t lim = 0xFFFFFFFFull;
  if (im.version != 1 || im.num_fields == 0 || im.num_fields > PB_MAX_FIELDS) return false;
  if (im.n_nodes == 0 || im.n_nodes > lim || im.n_edges > lim || im.n_terms > lim || im.n_docs > lim) return false;
  if (im.n_rows > lim || im.n_rows_padded > lim + 256 || im.n_rows_padded % 128 != 0 || im.n_rows_padded < im.n_rows + 128) return false;
  if (im.n_removed > im.n_docs || im.n_live_docs > im.n_docs) return false;
  return true;
}

}  // namespace

namespace pb {

// Structural validation of an image (the device kernels index with these values unchecked): offsets
// monotone and in range, children / terms / doc ordinals in range, every (tf, field length) within the
// header's max_tf / max_fl (pb_index_create sizes its u16 posting codes and BM25 table from them).
int validate_image(const pb_index_image* imp) {
  const pb_index_image& im = *imp;
  if (!scalars_sane(im)) { set_error("index image: header scalars out of range"); return PB_ERR_INVALID; }
  const uint64_t NN = im.n_nodes, NE = im.n_edges, NT = im.n_terms, R = im.n_rows, ND = im.n_docs, F = im.num_fields;
  if (!im.node_edge_begin || !im.node_term_lo || !im.node_term_hi || !im.node_parent || !im.node_char ||
      (NE && (!im.edge_char || !im.edge_child)) || !im.term_row_begin || (NT && (!im.term_byte_len || !im.term_node)) ||
      !im.post_blocks || (ND && !im.doc_key) || !im.removed_bitmap) {
    set_error("index image: null section"); return PB_ERR_INVALID;
  }
  if (im.node_edge_begin[0] != 0 || im.node_edge_begin[NN] != NE) { set_error("index image: node_edge_begin does not span the edges"); return PB_ERR_INVALID; }
  for (uint64_t n = 0; n < NN; ++n) {
    if (im.node_edge_begin[n + 1] < im.node_edge_begin[n]) { set_error("index image: node_edge_begin is not monotone at node %llu", (unsigned long long)n); return PB_ERR_INVALID; }
    if (im.node_term_lo[n] > im.node_term_hi[n] || im.node_term_hi[n] > NT) { set_error("index image: term range of node %llu out of range", (unsigned long long)n); return PB_ERR_INVALID; }
    if (n && im.node_parent[n] >= NN) { set_error("index image: parent of node %llu out of range", (unsigned long long)n); return PB_ERR_INVALID; }
    for (uint32_t e = im.node_edge_begin[n] + 1; e < im.node_edge_begin[n + 1]; ++e)
      if (im.edge_char[e] <= im.edge_char[e - 1]) { set_error("index image: edges of node %llu are not sorted by char", (unsigned long long)n); return PB_ERR_INVALID; }
  }
  for (uint64_t e = 0; e < NE; ++e)
    if (im.edge_child[e] >= NN) { set_error("index image: edge_child[%llu] out of range", (unsigned long long)e); return PB_ERR_INVALID; }
  if (im.term_row_begin[0] != 0 || im.term_row_begin[NT] != R) { set_error("index image: term_row_begin does not span the posting rows"); return PB_ERR_INVALID; }
  uint32_t max_bytes = 0;
  for (uint64_t t = 0; t < NT; ++t) {
    if (im.term_row_begin[t + 1] < im.term_row_begin[t]) { set_error("index image: term_row_begin is not monotone at term %llu", (unsigned long long)t); return PB_ERR_INVALID; }
    if (im.term_node[t] >= NN) { set_error("index image: term_node[%llu] out of range", (unsigned long long)t); return PB_ERR_INVALID; }
    max_bytes = im.term_byte_len[t] > max_bytes ? im.term_byte_len[t] : max_bytes;
  }
  if (max_bytes > im.max_term_bytes) { set_error("index image: a term is longer (%u bytes) than max_term_bytes (%u)", max_bytes, im.max_term_bytes); return PB_ERR_INVALID; }
  // posting columns: doc ordinals in range and ascending inside a term, (tf, fl) within the header's maxima
  const uint64_t NC = 1 + 2 * F;
  for (uint64_t t = 0; t < NT; ++t) {
    uint64_t prev = ~0ull;
    for (uint64_t r = im.term_row_begin[t]; r < im.term_row_begin[t + 1]; ++r) {
      const uint32_t* tile = im.post_blocks + (r / 128) * NC * 128;
      const uint32_t d = tile[r % 128];
      if (d >= ND || (prev != ~0ull && d <= prev)) { set_error("index image: doc ordinals of term %llu are out of range or not ascending", (unsigned long long)t); return PB_ERR_INVALID; }
      prev = d;
      for (uint64_t f = 0; f < F; ++f) {
        if (tile[(1 + f) * 128 + r % 128] > im.max_tf[f] || tile[(1 + F + f) * 128 + r % 128] > im.max_fl[f]) {
          set_error("index image: a (tf, field length) of term %llu exceeds the header's max_tf / max_fl", (unsigned long long)t); return PB_ERR_INVALID;
        }
      }
    }
  }
  uint64_t removed = 0;
  const uint64_t words = (ND + 31) / 32;
  for (uint64_t w = 0; w < words; ++w) removed += (uint64_t)__builtin_popcount(im.removed_bitmap[w]);
  if (ND % 32 && (im.removed_bitmap[words - 1] >> (ND % 32))) { set_error("index image: removed bits beyond n_docs"); return PB_ERR_INVALID; }
  if (removed != im.n_removed || im.n_live_docs + removed > ND) {
    set_error("index image: removed bitmap (%llu bits) disagrees with n_removed / n_live_docs", (unsigned long long)removed); return PB_ERR_INVALID;
  }
  return PB_OK;
}

}  // namespace pb

struct pb_image_file {
  std::vector<uint64_t> buf;      // 8-byte aligned storage of the whole file
  pb_index_image im{};
};

extern "C" {

int pb_image_save(const pb_index_image* im, const char* path) {
  if (!im || !path) { pb::set_error("pb_image_save: null argument"); return PB_ERR_INVALID; }
  if (!scalars_sane(*im)) { pb::set_error("pb_image_save: bad image header"); return PB_ERR_INVALID; }
  PB_TRY({
    Section sec[N_SECTIONS];
    sections_of(*im, sec);
    Scalars sc{};
    sc.version = im->version; sc.num_fields = im->num_fields;
    sc.n_nodes = im->n_nodes; sc.n_edges = im->n_edges; sc.n_terms = im->n_terms; sc.n_rows = im->n_rows;
    sc.n_rows_padded = im->n_rows_padded; sc.n_docs = im->n_docs; sc.max_term_bytes = im->max_term_bytes;
    for (uint32_t f = 0; f < PB_MAX_FIELDS; ++f) { sc.max_tf[f] = im->max_tf[f]; sc.max_fl[f] = im->max_fl[f]; sc.field_avg[f] = im->field_avg[f]; }
    sc.n_removed = im->n_removed; sc.n_live_docs = im->n_live_docs;
    const uint64_t header_bytes = align64(16 + sizeof(Scalars) + N_SECTIONS * 16 + 8);
    uint64_t table[N_SECTIONS][2];
    uint64_t off = header_bytes, sum = fnv1a(0xCBF29CE484222325ull, &sc, sizeof(sc));
    for (uint32_t i = 0; i < N_SECTIONS; ++i) {
      if (sec[i].bytes && !sec[i].p) { pb::set_error("pb_image_save: image section %u is null", i); return PB_ERR_INVALID; }
      table[i][0] = off; table[i][1] = sec[i].bytes;
      off = align64(off + sec[i].bytes);
      sum = fnv1a(sum, sec[i].p, sec[i].bytes);
    }
    std::unique_ptr<FILE, int (*)(FILE*)> f(std::fopen(path, "wb"), std::fclose);
    if (!f) { pb::set_error("pb_image_save: cannot open %s for writing", path); return PB_ERR_INVALID; }
    std::vector<uint8_t> head(header_bytes, 0);
    std::memcpy(head.data(), MAGIC, 8);
    const uint32_t hb = (uint32_t)header_bytes, ns = N_SECTIONS;
    std::memcpy(head.data() + 8, &hb, 4); std::memcpy(head.data() + 12, &ns, 4);
    std::memcpy(head.data() + 16, &sc, sizeof(sc));
    std::memcpy(head.data() + 16 + sizeof(sc), table, sizeof(table));
    std::memcpy(head.data() + 16 + sizeof(sc) + sizeof(table), &sum, 8);
    bool ok = std::fwrite(head.data(), 1, head.size(), f.get()) == head.size();
    static const uint8_t zeros[64] = {0};
    uint64_t pos = header_bytes;
    for (uint32_t i = 0; i < N_SECTIONS && ok; ++i) {
      if (sec[i].bytes) ok = std::fwrite(sec[i].p, 1, sec[i].bytes, f.get()) == sec[i].bytes;
      pos += sec[i].bytes;
      const uint64_t padn = align64(pos) - pos;
      if (ok && padn) ok = std::fwrite(zeros, 1, padn, f.get()) == padn;
      pos += padn;
    }
    if (!ok || std::fflush(f.get()) != 0) { pb::set_error("pb_image_save: short write to %s", path); return PB_ERR_INVALID; }
    return PB_OK;
  });
}

int pb_image_load(const char* path, pb_image_file** out) {
  if (!path || !out) { pb::set_error("pb_image_load: null argument"); return PB_ERR_INVALID; }
  PB_TRY({
    std::unique_ptr<FILE, int (*)(FILE*)> f(std::fopen(path, "rb"), std::fclose);
    if (!f) { pb::set_error("pb_image_load: cannot open %s", path); return PB_ERR_INVALID; }
    std::fseek(f.get(), 0, SEEK_END);
    const long size = std::ftell(f.get());
    std::fseek(f.get(), 0, SEEK_SET);
    if (size < 64) { pb::set_error("pb_image_load: %s is not an index image (too short)", path); return PB_ERR_INVALID; }
    std::unique_ptr<pb_image_file> h(new pb_image_file());
    h->buf.resize(((uint64_t)size + 7) / 8 + 8, 0);
    uint8_t* base = reinterpret_cast<uint8_t*>(h->buf.data());
    if (std::fread(base, 1, (size_t)size, f.get()) != (size_t)size) { pb::set_error("pb_image_load: short read from %s", path); return PB_ERR_INVALID; }
    if (std::memcmp(base, MAGIC, 8) != 0) { pb::set_error("pb_image_load: %s is not an index image (bad magic)", path); return PB_ERR_INVALID; }
    uint32_t hb = 0, ns = 0;
    std::memcpy(&hb, base + 8, 4); std::memcpy(&ns, base + 12, 4);
    if (ns != N_SECTIONS || hb != align64(16 + sizeof(Scalars) + N_SECTIONS * 16 + 8) || hb > (uint64_t)size) {
      pb::set_error("pb_image_load: %s has an unknown header layout", path); return PB_ERR_INVALID;
    }
    Scalars sc;
    std::memcpy(&sc, base + 16, sizeof(sc));
    uint64_t table[N_SECTIONS][2];
    std::memcpy(table, base + 16 + sizeof(sc), sizeof(table));
    uint64_t want_sum = 0;
    std::memcpy(&want_sum, base + 16 + sizeof(sc) + sizeof(table), 8);
    if (sc.version != 1 || sc.num_fields == 0 || sc.num_fields > PB_MAX_FIELDS) { pb::set_error("pb_image_load: bad image header in %s", path); return PB_ERR_INVALID; }
    pb_index_image& im = h->im;
    im.version = sc.version; im.num_fields = sc.num_fields;
    im.n_nodes = sc.n_nodes; im.n_edges = sc.n_edges; im.n_terms = sc.n_terms; im.n_rows = sc.n_rows;
    im.n_rows_padded = sc.n_rows_padded; im.n_docs = sc.n_docs; im.max_term_bytes = sc.max_term_bytes;
    for (uint32_t x = 0; x < PB_MAX_FIELDS; ++x) { im.max_tf[x] = sc.max_tf[x]; im.max_fl[x] = sc.max_fl[x]; im.field_avg[x] = sc.field_avg[x]; }
    im.n_removed = sc.n_removed; im.n_live_docs = sc.n_live_docs;
    if (!scalars_sane(im)) { pb::set_error("pb_image_load: header scalars of %s are out of range", path); return PB_ERR_INVALID; }
    // the section sizes the scalars imply must be the sizes on file
    Section expect[N_SECTIONS];
    sections_of(im, expect);
    uint64_t sum = fnv1a(0xCBF29CE484222325ull, &sc, sizeof(sc));
    const void* ptr[N_SECTIONS];
    for (uint32_t i = 0; i < N_SECTIONS; ++i) {
      const uint64_t off = table[i][0], n = table[i][1];
      if (n != expect[i].bytes || (off & 63) || off < hb || off > (uint64_t)size || n > (uint64_t)size - off) {
        pb::set_error("pb_image_load: section %u of %s is inconsistent with the header", i, path); return PB_ERR_INVALID;
      }
      ptr[i] = base + off;
      sum = fnv1a(sum, ptr[i], n);
    }
    if (sum != want_sum) { pb::set_error("pb_image_load: checksum mismatch in %s (file is corrupt)", path); return PB_ERR_INVALID; }
    im.node_edge_begin = (const uint32_t*)ptr[0]; im.node_term_lo = (const uint32_t*)ptr[1]; im.node_term_hi = (const uint32_t*)ptr[2];
    im.node_parent = (const uint32_t*)ptr[3]; im.node_char = (const uint32_t*)ptr[4]; im.edge_char = (const uint32_t*)ptr[5];
    im.edge_child = (const uint32_t*)ptr[6]; im.term_row_begin = (const uint64_t*)ptr[7]; im.term_byte_len = (const uint32_t*)ptr[8];
    im.term_node = (const uint32_t*)ptr[9]; im.post_blocks = (const uint32_t*)ptr[10]; im.doc_key = (const uint64_t*)ptr[11];
    im.removed_bitmap = (const uint32_t*)ptr[12];
    {
      const int rc = pb::validate_image(&im);
      if (rc != PB_OK) return rc;
    }
    *out = h.release();
    return PB_OK;
  });
}

const pb_index_image* pb_image_file_image(const pb_image_file* f) { return f ? &f->im : nullptr; }

void pb_image_file_free(pb_image_file* f) { delete f; }

}  // extern "C"
