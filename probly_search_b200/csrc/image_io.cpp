// On-disk / wire format of the flattened index image (the reference has no
// serialisation at all — no serde, the index lives only in RAM).  One file = the pb_index_image a
// builder flattens: header, section table, 64-byte aligned sections, FNV-1a checksum of the payload.
// A loaded file yields a pb_index_image whose pointers point into the file's buffer, ready for
// pb_index_create: a process can serve queries without ever holding the mutable host index.
//
//   offset 0   magic "PBIMG1\0\0"            8 bytes
//          8   header_bytes (u32)  n_sections (u32)
//         16   scalar block: the non-pointer fields of pb_index_image, little endian, in order
//              (version, num_fields, n_nodes, n_edges, n_terms, n_rows, n_rows_padded, n_docs,
//               max_term_bytes, max_tf[4], max_fl[4], n_removed, n_live_docs, field_avg[4])
//          ..  section table: n_sections x {offset u64, bytes u64}
//          ..  payload_checksum (u64, FNV-1a 64 over all section bytes in table order)
//          ..  sections, each starting on a 64-byte boundary
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

#include "../../include/probly_b200.h"
#include "common.hpp"

namespace {

constexpr char MAGIC[8] = {'P', 'B', 'I', 'M', 'G', '1', 0, 0};
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

}  // namespace

struct pb_image_file {
  std::vector<uint64_t> buf;      // 8-byte aligned storage of the whole file
  pb_index_image im{};
};

extern "C" {

int pb_image_save(const pb_index_image* im, const char* path) {
  if (!im || !path) { pb::set_error("pb_image_save: null argument"); return PB_ERR_INVALID; }
  if (im->version != 1 || im->num_fields == 0 || im->num_fields > PB_MAX_FIELDS) { pb::set_error("pb_image_save: bad image header"); return PB_ERR_INVALID; }
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
    uint64_t off = header_bytes, sum = 0xCBF29CE484222325ull;
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
    // the section sizes the scalars imply must be the sizes on file
    Section expect[N_SECTIONS];
    sections_of(im, expect);
    uint64_t sum = 0xCBF29CE484222325ull;
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
    *out = h.release();
    return PB_OK;
  });
}

const pb_index_image* pb_image_file_image(const pb_image_file* f) { return f ? &f->im : nullptr; }

void pb_image_file_free(pb_image_file* f) { delete f; }

}  // extern "C"
