// Host orchestration of the device query path + the pb_index / pb_batch half of the C ABI
// (include/probly_b200.h).  One pb_index = one flattened image resident in HBM on one device;
// one pb_batch = one uploaded query batch, its stream and its workspace.
//
// pb_batch_run pipeline (all on the batch's stream):
//   descend -> plan (classify, segment descriptors, tile prefix sums) -> score kernel over all
//   single-list queries (the dominant launch) -> rounds of the multi-list side path
//   (mark / score+divert / radix sort / fold / clear) -> finalize (merge partial top-k lists).
// There is no CPU fallback anywhere in this file: without a device every entry point fails.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <atomic>
#include <condition_variable>
#include <thread>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <type_traits>
#include <vector>

#include <cub/cub.cuh>

#include "common.hpp"
#include "field_ops.hpp"
#include "group.hpp"
#include "plan_kernels.cuh"

using namespace pbk;
typedef unsigned long long ull;

#define CU(x)                                                                                    \
  do {                                                                                           \
    cudaError_t e_ = (x);                                                                        \
    if (e_ != cudaSuccess) {                                                                     \
      pb::set_error("%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__);    \
      return (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver) ? PB_ERR_NO_DEVICE   \
                                                                            : PB_ERR_CUDA;       \
    }                                                                                            \
  } while (0)
#define RC(x)                  \
  do {                         \
    int rc_ = (x);             \
    if (rc_ != PB_OK) return rc_; \
  } while (0)

namespace {

template <class T>
struct DBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t c = n + n / 8 + 64;
    cudaError_t e = cudaMalloc((void**)&p, c * sizeof(T));
    if (e == cudaSuccess) {
      // defined contents (slack and unused slots get copied / scanned).  The fill runs on the default stream while the
      // batches use non-blocking streams: wait for it, or it could land AFTER an upload into the new buffer.
      cap = c;
      e = cudaMemset(p, 0, c * sizeof(T));
      if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    }
    return e;
  }
  // grow, keeping the first `keep` elements (device-to-device copy on `st`)
  cudaError_t ensure_keep(size_t n, size_t keep, cudaStream_t st) {
    if (n <= cap && p) return cudaSuccess;
    size_t c = n + n / 2 + 64;
    T* np_ = nullptr;
    cudaError_t e = cudaMalloc((void**)&np_, c * sizeof(T));
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(np_, 0, c * sizeof(T), st);
    if (e != cudaSuccess) { cudaFree(np_); return e; }
    if (p && keep) e = cudaMemcpyAsync(np_, p, std::min(keep, cap) * sizeof(T), cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (p) cudaFree(p);
    p = np_; cap = c;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  ~DBuf() { release(); }
  DBuf() = default;
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
};

template <class T>
cudaError_t upload(DBuf<T>& d, const T* h, size_t n, size_t extra_zero = 0) {
  cudaError_t e = d.ensure(n + extra_zero + 1);
  if (e != cudaSuccess) return e;
  if (n) { e = cudaMemcpy(d.p, h, n * sizeof(T), cudaMemcpyHostToDevice); if (e != cudaSuccess) return e; }
  return cudaMemset(d.p + n, 0, (extra_zero + 1) * sizeof(T));
}

uint32_t bits_for(uint64_t n) {   // bits needed to represent values 0..n-1 (at least 1)
  uint32_t b = 1;
  while ((1ull << b) < n) ++b;
  return b;
}

int g_sm_count(int device) {
  int n = 148;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
  return n;
}

}  // namespace

// ==========================================================================================
// pb_index
// ==========================================================================================
struct pb_index {
  int device = 0;
  int sm_count = 148;
  uint32_t F = 1;
  uint64_t n_nodes = 0, n_edges = 0, n_terms = 0, n_rows = 0, n_rows_padded = 0, n_docs = 0;
  uint32_t max_term_bytes = 0, max_tf[4] = {0, 0, 0, 0}, max_fl[4] = {0, 0, 0, 0};
  bool narrow = false;          // device posting columns hold one u16 (tf, fl) code per field (see IndexView)
  uint32_t tile_words = 0, fl_bits[4] = {0, 0, 0, 0};
  DBuf<uint32_t> node_edge_begin, node_term_lo, node_term_hi, edge_char, edge_child;
  DBuf<uint64_t> term_row_begin;
  DBuf<uint32_t> term_byte_len, post_blocks, removed, live_prefix, term_live_rows;
  DBuf<uint64_t> term_df_live, liverows_prefix;
  DBuf<double> term_idf, eb;
  DBuf<double> rcp;             // RN(1 / d), d = 0..1024 (PB_Z2O_RCP experiment, kernels.cuh)
  DBuf<uint2> dir;              // rank directories of the dense posting lists (IndexView::dir)
  DBuf<uint32_t> term_dir;
  uint32_t dir_words = 0, n_dense = 0;
  // union image (union_kernels.cuh): the posting rows re-sorted by (doc shard, term, doc)
  DBuf<uint32_t> u_meta, u_term, u_shard_row;
  DBuf<uint16_t> u_code[4];
  uint32_t u_wbits = 0, u_shards = 0;
  bool u_ok = false;
  bool u_warp = false;           // union_warp_kernel (small shards, one warp per task) instead of union_kernel (2048-doc shards per CTA)
  DBuf<ull> liverowcnt_prefix, dflive_prefix;   // per-term prefixes of term_live_rows / term_df_live (class-U row statistics)
  DBuf<uint32_t> row_dead;       // IndexView::row_dead: the removed bitmap per posting row (rebuilt with the live state)
  bool row_dead_ok = false;
  // compact copy of the narrow tiles (IndexView::cpost): u16 doc offsets, streamed by the single-list launch
  DBuf<uint32_t> cpost, cbase;
  DBuf<uint8_t> term_compact;
  bool compact = false;
  uint64_t compact_rows = 0;     // rows of the lists that stream from the compact copy
  // host copies needed to rebuild term strings (pb_index_expand_term) and to recompute idf
  std::vector<uint32_t> h_node_parent, h_node_char, h_term_node;
  std::vector<uint64_t> h_term_row_begin, h_df_live, h_df_extra;
  uint64_t n_live = 0, n_removed = 0;
  double avg[4] = {0, 0, 0, 0};
  uint64_t live_epoch = 0;       // bumped by every pb_index_set_live_state: staged batches rebuild their BM25 table
  std::mutex occ_mu;             // guards occ_cache
  std::map<uint64_t, int> occ_cache;   // resident CTAs per SM of a scoring-kernel shape
  // delta segment (pb_index_attach_delta): the rows of the documents added behind this image, under the current trie
  pb_index* delta = nullptr;     // owned
  std::vector<uint32_t> sid;     // builder term id of every term ordinal of THIS image (matches terms across the segments)
  // pb_query_batch / pb_query_full / pb_index_expand_term borrow one of a few internal batches (own stream + workspace
  // each), so calls from different host threads overlap on the device; they hold state_mu shared, the calls that
  // change the index (pb_index_set_live_state, pb_index_set_df_extra, pb_index_attach_delta) hold it exclusively.
  std::shared_mutex state_mu;
  std::atomic<int> writers_waiting{0};   // one-call queries stand back while a state change waits (no writer starvation)
  std::mutex pool_mu;            // guards the four members below
  std::condition_variable pool_cv;
  std::vector<pb_batch*> pool_free;
  int pool_size = 0;
  pb_batch_stats last_st{};      // stats of the most recently finished one-call query (pb_index_last_stats)
  bool has_last = false;

  UnionView uview() const {
    UnionView u;
    u.meta = u_meta.p; u.term = u_term.p;
    for (int f = 0; f < 4; ++f) u.code[f] = u_code[f].p;
    u.shard_row = u_shard_row.p; u.n_shards = u_shards; u.wbits = u_wbits; u.ok = u_ok ? 1u : 0u;
    return u;
  }
  IndexView view() const {
    IndexView v;
    v.node_edge_begin = node_edge_begin.p; v.node_term_lo = node_term_lo.p; v.node_term_hi = node_term_hi.p;
    v.edge_char = edge_char.p; v.edge_child = edge_child.p;
    v.term_row_begin = term_row_begin.p; v.term_byte_len = term_byte_len.p;
    v.post_blocks = post_blocks.p; v.tile_words = tile_words; v.narrow = narrow ? 1u : 0u;
    for (int f = 0; f < 4; ++f) v.fl_bits[f] = fl_bits[f];
    v.cpost = compact ? cpost.p : nullptr; v.cbase = compact ? cbase.p : nullptr; v.term_compact = compact ? term_compact.p : nullptr;
    v.removed = removed.p;
    v.row_dead = (n_removed && row_dead_ok) ? row_dead.p : nullptr;
    v.term_df_live = term_df_live.p; v.term_live_rows = term_live_rows.p; v.live_prefix = live_prefix.p; v.liverows_prefix = liverows_prefix.p;
    v.term_idf = term_idf.p; v.eb = eb.p;
    v.dir = dir.p; v.term_dir = term_dir.p; v.dir_words = dir_words;
    v.rcp = rcp.p; v.rcp_ok = max_term_bytes <= 255u ? 1u : 0u;
    v.n_terms = (uint32_t)n_terms; v.n_docs = (uint32_t)n_docs; v.num_fields = F;
    v.has_removed = n_removed ? 1u : 0u;
    return v;
  }
};

// Compact copy of the narrow tiles (IndexView::cpost), built on the device from the tiles already in
// HBM — only on request (PB_POSTING_COMPACT=1).  Measured on B200 ("what was tried"): the single-list
// stream reads 25 % fewer bytes from it and is 5-6 % SLOWER, on the 2.1 GB image of cfg 3/4 as on the L2-resident
// one of cfg 1: the stream is bound by instruction issue / dependent-issue latency, and rebuilding the doc ordinals
// costs two instructions per row.  Kept as a tested layout for bandwidth-starved parts, not used by default.
static int index_build_compact(pb_index* ix) {
  ix->compact = false;
  ix->compact_rows = 0;
  if (!ix->narrow || ix->n_rows < (uint64_t)TILE_ROWS || ix->n_terms == 0) return PB_OK;
  const uint64_t tiles = ix->n_rows_padded / TILE_ROWS;
  const char* e = std::getenv("PB_POSTING_COMPACT");
  if (!(e && !std::strcmp(e, "1"))) return PB_OK;
  const uint32_t CW = (TILE_ROWS / 2) * (1 + ix->F);
  const uint64_t alloc_tiles = tiles + 2;                        // the streaming loop loads one tile ahead
  CU(ix->cpost.ensure(alloc_tiles * CW + 64));                   // zero-filled
  CU(ix->cbase.ensure(alloc_tiles + 2));
  CU(ix->term_compact.ensure(ix->n_terms + 1));
  DBuf<ull> d_rows;
  CU(d_rows.ensure(1));
  compact_tiles_kernel<<<ix->sm_count * 8, 256>>>(ix->post_blocks.p, ix->tile_words, ix->F, ix->n_rows / TILE_ROWS, alloc_tiles + 2,
                                                  ix->cpost.p, ix->cbase.p);
  CU(cudaGetLastError());
  uint32_t min_tiles = 4;                                        // shorter lists are all edges and set-up anyway
  if (const char* e = std::getenv("PB_COMPACT_MIN_TILES")) min_tiles = (uint32_t)std::max(1, atoi(e));   // tests
  term_compact_kernel<<<ix->sm_count * 8, 256>>>(ix->term_row_begin.p, (uint32_t)ix->n_terms, ix->cbase.p, min_tiles,
                                                 ix->term_compact.p, d_rows.p);
  CU(cudaGetLastError());
  ull h_rows = 0;
  CU(cudaMemcpy(&h_rows, d_rows.p, sizeof(ull), cudaMemcpyDeviceToHost));
  ix->compact_rows = h_rows;
  ix->compact = h_rows > 0;
  if (!ix->compact) { ix->cpost.release(); ix->cbase.release(); ix->term_compact.release(); }
  return PB_OK;
}

// Builds the union image on the device from the term-major image already resident in HBM: a stable
// radix sort of (doc shard, row) keeps the rows of a shard in (term, doc) order.
static int index_build_union(pb_index* ix) {
  ix->u_ok = false;
  const char* e = std::getenv("PB_UNION");
  if (e && !std::strcmp(e, "0")) return PB_OK;
  bool ok = ix->max_term_bytes <= 63u && ix->n_rows > 0 && ix->n_rows < 0xFFFFFF00ull && ix->n_terms > 0;
  for (uint32_t f = 0; f < ix->F; ++f) {
    ok = ok && ix->max_tf[f] <= 63u && ix->fl_bits[f] <= 8u;
    ok = ok && (((uint64_t)ix->max_tf[f] + 1) << ix->fl_bits[f]) <= 65536ull;
  }
  if (!ok) return PB_OK;                      // outside the envelope of the dense path: queries take the per-list path
  const uint64_t R = ix->n_rows, RP = R + 256;          // pad: 128-row groups are loaded whole
  // two kernels: union_warp_kernel owns a 512-doc shard per WARP, union_kernel a 2048-doc shard per CTA (PB_UNION_KERNEL=cta)
  {
    const char* k = std::getenv("PB_UNION_KERNEL");
    ix->u_warp = !(k && !std::strcmp(k, "cta"));
  }
  ix->u_wbits = ix->u_warp ? 9 : 11;
  if (const char* w = std::getenv("PB_UNION_WBITS")) ix->u_wbits = (uint32_t)std::min(ix->u_warp ? 10 : 11, std::max(ix->u_warp ? 7 : 11, atoi(w)));   // tuning knob
  ix->u_shards = (uint32_t)(((ix->n_docs ? ix->n_docs - 1 : 0) >> ix->u_wbits) + 1);
  DBuf<uint32_t> keys, vals, keys2, vals2, row_term;
  DBuf<uint8_t> tmp;
  CU(keys.ensure(R)); CU(vals.ensure(R)); CU(keys2.ensure(R)); CU(vals2.ensure(R)); CU(row_term.ensure(R));
  IndexView v = ix->view();
  ubuild_keys_kernel<<<ix->sm_count * 8, 256>>>(v, ix->u_wbits, keys.p, vals.p, row_term.p);
  CU(cudaGetLastError());
  const int end_bit = (int)bits_for((uint64_t)ix->u_shards + 1);
  size_t bytes = 0;
  CU(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys.p, keys2.p, vals.p, vals2.p, (int64_t)R, 0, end_bit));
  CU(tmp.ensure(bytes));
  bytes = tmp.cap;
  CU(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, keys.p, keys2.p, vals.p, vals2.p, (int64_t)R, 0, end_bit));
  CU(ix->u_meta.ensure(RP)); CU(ix->u_term.ensure(RP));
  CU(cudaMemset(ix->u_meta.p, 0, ix->u_meta.cap * sizeof(uint32_t)));
  CU(cudaMemset(ix->u_term.p, 0xFF, ix->u_term.cap * sizeof(uint32_t)));
  for (uint32_t f = 0; f < ix->F; ++f) {
    CU(ix->u_code[f].ensure(RP));
    CU(cudaMemset(ix->u_code[f].p, 0, ix->u_code[f].cap * sizeof(uint16_t)));
  }
  ubuild_gather_kernel<<<(unsigned)((R + 255) / 256), 256>>>(v, ix->u_wbits, R, vals2.p, row_term.p, ix->u_meta.p, ix->u_term.p,
                                                             ix->u_code[0].p, ix->u_code[1].p, ix->u_code[2].p, ix->u_code[3].p);
  CU(cudaGetLastError());
  CU(ix->u_shard_row.ensure(ix->u_shards + 2));
  ubuild_bounds_kernel<<<(ix->u_shards + 1 + 255) / 256, 256>>>(keys2.p, R, ix->u_shards, ix->u_shard_row.p);
  CU(cudaGetLastError());
  CU(cudaDeviceSynchronize());
  ix->u_ok = true;
  return PB_OK;
}

// BM25::before_each, bm25.rs:41-56: idf from the LIVE doc count and the clamped live df.
// libm `log` on the host = what Rust's f64::ln calls, so the table is bit-identical.
// h_df_extra (pb_index_set_df_extra): live occurrences of the same term in the OTHER segments of a segmented index —
// the reference counts one posting list per term, so the document frequency is the sum over the segments.
static int index_upload_idf(pb_index* ix) {
  const size_t NT = ix->n_terms;
  const bool extra = ix->h_df_extra.size() == NT;
  std::vector<double> idf(NT + 1, 0.0);
  for (size_t t = 0; t < NT; ++t) {
    const uint64_t df = ix->h_df_live[t] + (extra ? ix->h_df_extra[t] : 0);
    const uint64_t frequency = std::min<uint64_t>(ix->n_live, df);
    const uint64_t diff = ix->n_live - frequency;
    idf[t] = std::log(1.0 + ((double)diff + 0.5) / ((double)frequency + 0.5));
  }
  CU(ix->term_idf.ensure(NT + 1));
  CU(cudaMemcpy(ix->term_idf.p, idf.data(), (NT + 1) * sizeof(double), cudaMemcpyHostToDevice));
  CU(cudaDeviceSynchronize());     // see index_apply_live_state: the DMA of a pageable copy may still be in flight
  return PB_OK;
}

static int index_apply_live_state(pb_index* ix, const uint32_t* bitmap_words, uint64_t n_removed,
                                  uint64_t n_live, const double* avg) {
  CU(cudaSetDevice(ix->device));
  const size_t words = (ix->n_docs + 31) / 32 + 1;
  CU(ix->removed.ensure(words));
  CU(cudaMemcpy(ix->removed.p, bitmap_words, words * sizeof(uint32_t), cudaMemcpyHostToDevice));
  ix->n_removed = n_removed;
  ix->n_live = n_live;
  ++ix->live_epoch;
  // the same bitmap per posting row, for the scoring loop (PB_ROW_DEAD=0: keep probing the doc bitmap - A/B and tests)
  ix->row_dead_ok = false;
  if (n_removed && ix->n_rows && ix->narrow) {          // the wide layout's loop probes the bitmap (process_tile)
    const char* e = std::getenv("PB_ROW_DEAD");
    if (!(e && !std::strcmp(e, "0"))) {
      const uint64_t tiles = ix->n_rows_padded / TILE_ROWS;
      CU(ix->row_dead.ensure((tiles + 3) * 4));                  // zero-filled; + the look-ahead tiles
      row_dead_kernel<<<ix->sm_count * 8, 256>>>(ix->post_blocks.p, ix->tile_words, tiles, ix->removed.p, (uint32_t)ix->n_docs, ix->row_dead.p);
      CU(cudaGetLastError());
      CU(cudaDeviceSynchronize());
      ix->row_dead_ok = true;
    }
  }
  for (uint32_t f = 0; f < ix->F; ++f) ix->avg[f] = avg[f];
  const size_t NT = ix->n_terms;
  CU(ix->term_df_live.ensure(NT + 1));
  CU(ix->term_live_rows.ensure(NT + 1));
  CU(cudaMemset(ix->term_live_rows.p, 0, (NT + 1) * sizeof(uint32_t)));
  CU(ix->live_prefix.ensure(NT + 2));
  CU(ix->liverows_prefix.ensure(NT + 2));
  CU(ix->term_idf.ensure(NT + 1));
  CU(cudaMemset(ix->term_df_live.p, 0, (NT + 1) * sizeof(uint64_t)));
  if (NT) {
    IndexView v = ix->view();
    int grid = ix->sm_count * 8;
    CU(field_ops(ix->F)->live_df_launch(&v, (ull*)ix->term_df_live.p, ix->term_live_rows.p, grid, 0));
  }
  ix->h_df_live.assign(NT + 1, 0);
  CU(cudaMemcpy(ix->h_df_live.data(), ix->term_df_live.p, NT * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  std::vector<uint32_t> lp(NT + 2, 0);
  std::vector<uint64_t> lrp(NT + 2, 0);
  for (size_t t = 0; t < NT; ++t) {
    uint64_t df = ix->h_df_live[t];
    lp[t + 1] = lp[t] + (df > 0 ? 1u : 0u);
    lrp[t + 1] = lrp[t] + (df > 0 ? (ix->h_term_row_begin[t + 1] - ix->h_term_row_begin[t]) : 0);
  }
  {
    std::vector<uint32_t> h_live_rows(NT + 1, 0);
    if (NT) CU(cudaMemcpy(h_live_rows.data(), ix->term_live_rows.p, NT * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    std::vector<ull> lrc(NT + 2, 0), dfp(NT + 2, 0);
    for (size_t t = 0; t < NT; ++t) { lrc[t + 1] = lrc[t] + h_live_rows[t]; dfp[t + 1] = dfp[t] + ix->h_df_live[t]; }
    CU(ix->liverowcnt_prefix.ensure(NT + 2)); CU(ix->dflive_prefix.ensure(NT + 2));
    CU(cudaMemcpy(ix->liverowcnt_prefix.p, lrc.data(), (NT + 1) * sizeof(ull), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ix->dflive_prefix.p, dfp.data(), (NT + 1) * sizeof(ull), cudaMemcpyHostToDevice));
  }
  RC(index_upload_idf(ix));
  CU(cudaMemcpy(ix->live_prefix.p, lp.data(), (NT + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(ix->liverows_prefix.p, lrp.data(), (NT + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice));
  // a cudaMemcpy from pageable memory may return before its DMA has landed; the batches run on non-blocking streams
  CU(cudaDeviceSynchronize());
  return PB_OK;
}

// ==========================================================================================
// pb_batch
// ==========================================================================================
struct pb_batch {
  pb_index* ix = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[9] = {};
  uint64_t Q = 0, NT = 0;
  uint32_t scorer = 0, k = 0;
  double k1 = 1.2, b = 0.75, boost[4] = {1, 1, 1, 1};
  bool loaded = false, ran = false;
  // capture of full result sets
  uint64_t full_cap = 0;
  DBuf<uint32_t> full_q, full_doc;
  DBuf<double> full_score;
  // inputs
  DBuf<uint8_t> term_bytes;
  DBuf<uint64_t> term_byte_off, query_term_off;
  // plan
  DBuf<uint32_t> qt_lo, qt_hi, qt_len, qt_q;
  DBuf<ull> qt_gcount, qt_goff, q_isg, q_gidx, q_grows, q_prim, q_recbound, q_recoff, q_nbins, q_binoff, q_gsegoff, q_gtileoff, q_bmwords, q_bmoff;
  DBuf<uint8_t> q_scheme, q_shift;
  DBuf<uint32_t> bin_count, bin_off, bin_cursor;
  DBuf<uint4> rec;
  DBuf<ull> s_tiles, s_tile_off, g_tiles, g_tile_off, g_mtiles, g_mtile_off;
  DBuf<Seg> seg_s, seg_g;
  DBuf<uint8_t> cub_temp;
  // outputs: ONE packed block laid out for `slot` >= Q queries
  //   [n_results u64 x slot][doc_digest u64 x slot][score_digest u64 x slot][topk_score f64 x slot*k]
  //   [topk_doc u32 x slot*k][topk_n u32 x slot]   (padded to 16 bytes)
  // so that the multi-GPU exchange is a single ncclAllGather of the block.
  DBuf<uint8_t> res, gath;
  uint64_t slot = 0;            // queries the block is laid out for (= Q unless a gather asks for more)
  uint64_t want_slot = 0;       // pb_batch_set_gather: the per-rank slot every rank agreed on
  size_t block_bytes = 0;
  ull *n_results = nullptr, *doc_digest = nullptr, *score_digest = nullptr;
  uint32_t *topk_n = nullptr, *topk_doc = nullptr;
  double* topk_score = nullptr;
  pb_comm* comm = nullptr;      // not owned
  bool gather_pending = false, gathered = false;
  uint64_t tab_epoch = 0;       // ix->live_epoch the BM25 table was built for
  // partial lists + counters
  DBuf<uint32_t> part_head, part_next, part_n, part_doc;
  DBuf<double> part_score;
  // every small counter of a run lives in ONE arena (one memset per run, one D2H for the stats):
  //   counters [8 x u32]: [0] part_count [1] rec_count [2] error [3] overflow   stats [3][ST_COUNT]: S, G, U phase
  //   u_counter [8]: union work-item counters + cycle profile   full_count [2]   xcount [4]
  template <class T> struct View { T* p = nullptr; };
  DBuf<ull> scal;
  static constexpr size_t SCAL_WORDS = 4 + 3 * ST_COUNT + 8 + 2 + 4;
  View<uint32_t> counters;
  View<ull> stats, u_counter, full_count, xcount;
  void* h_stage = nullptr;     // pinned staging for small result blocks (pb_query_batch: one D2H instead of six)
  size_t h_stage_bytes = 0;
  bool tab_valid = false;      // the BM25 table on the device matches (tab_k1, tab_b, tab_epoch)
  double tab_k1 = 0, tab_b = 0;
  double tab_scale[4] = {1.0, 1.0, 1.0, 1.0};   // power-of-two boosts folded into the resident table (ScoreParams::tab_scale)
  DBuf<UQuery> uq;
  DBuf<ull> q_isu, q_uidx, q_isu2, q_uidx2;
  DBuf<uint32_t> u_list, u_list2;
  // side path
  DBuf<uint32_t> bitmap;
  size_t bitmap_zeroed = 0;
  DBuf<ull> rec_key, rec_val, rec_key2, rec_val2;
  DBuf<uint8_t> rec_flags;
  // BM25 table
  DBuf<double> tab;
  uint32_t tab_tfcap[4] = {}, tab_flcap[4] = {}, tab_off[4] = {}, tab_total = 0;
  bool tab_full = false;
  // host staging
  std::vector<ull> h_recoff, h_binoff, h_gidx, h_gsegoff, h_gtileoff, h_bmoff;
  bool h_full = false;           // the vectors above hold every entry (else only [0] and [Q])
  pb_batch_stats st{};
  uint32_t launches = 0;         // kernels enqueued by the current run
  // per-launch-class timing inside the side path: pairs of pooled events, summed by batch_finish
  std::vector<cudaEvent_t> rev;
  size_t rev_used = 0;
  std::vector<std::pair<uint32_t, uint32_t>> rev_span[4];   // 0 mark, 1 score (class G), 2 fold (bin fold + sorted fallback), 3 union
  int rev_begin() {
    if (rev_used == rev.size()) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return -1; rev.push_back(e); }
    if (cudaEventRecord(rev[rev_used], stream) != cudaSuccess) return -1;
    return (int)rev_used++;
  }
  void rev_end(int cls, int e0) {
    int e1 = rev_begin();
    if (e0 >= 0 && e1 >= 0) rev_span[cls].push_back({(uint32_t)e0, (uint32_t)e1});
  }

  ~pb_batch() {
    if (ix) cudaSetDevice(ix->device);
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    for (auto& e : rev) cudaEventDestroy(e);
    if (h_stage) cudaFreeHost(h_stage);
    if (stream) cudaStreamDestroy(stream);
  }
};

namespace {

int scan_ull(pb_batch* b, const ull* in, ull* out, size_t n_plus_1) {
  size_t bytes = 0;
  CU(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int64_t)n_plus_1, b->stream));
  CU(b->cub_temp.ensure(bytes));
  bytes = b->cub_temp.cap;
  CU(cub::DeviceScan::ExclusiveSum(b->cub_temp.p, bytes, in, out, (int64_t)n_plus_1, b->stream));
  return PB_OK;
}

__global__ void overflow_count_kernel(const uint32_t* __restrict__ bin_cursor, uint32_t n_bins, uint32_t* __restrict__ out) {
  uint32_t s = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_bins; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t c = bin_cursor[i];
    if (c > PB_WINDOW_MAX) s += c;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

int scan_u32(pb_batch* b, const uint32_t* in, uint32_t* out, size_t n_plus_1) {
  size_t bytes = 0;
  CU(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int64_t)n_plus_1, b->stream));
  CU(b->cub_temp.ensure(bytes));
  bytes = b->cub_temp.cap;
  CU(cub::DeviceScan::ExclusiveSum(b->cub_temp.p, bytes, in, out, (int64_t)n_plus_1, b->stream));
  return PB_OK;
}

__global__ void gather_tileoff_kernel(uint64_t n, const ull* __restrict__ q_gsegoff,
                                      const ull* __restrict__ g_tile_off, ull* __restrict__ q_gtileoff) {
  uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (q < n) q_gtileoff[q] = g_tile_off[q_gsegoff[q]];
}

struct ResLayout { size_t n, dd, sd, ts, td, tn, bytes; };
ResLayout res_layout(uint64_t slot, uint32_t k) {
  const size_t kk = std::max<uint32_t>(k, 1);
  ResLayout L;
  L.n = 0; L.dd = L.n + slot * 8; L.sd = L.dd + slot * 8; L.ts = L.sd + slot * 8;
  L.td = L.ts + slot * kk * 8; L.tn = L.td + slot * kk * 4;
  L.bytes = (L.tn + slot * 4 + 15) & ~(size_t)15;
  return L;
}
// (re)lays the packed result block out for max(Q, want_slot) queries
int batch_layout_results(pb_batch* b) {
  const uint64_t slot = std::max<uint64_t>(std::max<uint64_t>(b->Q, b->want_slot), 1);
  const ResLayout L = res_layout(slot, b->k);
  CU(b->res.ensure(L.bytes));
  b->slot = slot; b->block_bytes = L.bytes;
  uint8_t* p = b->res.p;
  b->n_results = (ull*)(p + L.n); b->doc_digest = (ull*)(p + L.dd); b->score_digest = (ull*)(p + L.sd);
  b->topk_score = (double*)(p + L.ts); b->topk_doc = (uint32_t*)(p + L.td); b->topk_n = (uint32_t*)(p + L.tn);
  if (b->comm) CU(b->gath.ensure(L.bytes * (size_t)pbg::comm_world(b->comm)));
  return PB_OK;
}

double bm25_tf_host(double k1, double b, double avg, uint32_t tf, uint32_t fl) {
  // bm25.rs:78-82 — same operation order; this file is compiled with -ffp-contract=off
  double tfd = (double)tf;
  return ((k1 + 1.0) * tfd) / (k1 * ((1.0 - b) + b * ((double)fl / avg)) + tfd);
}

// The factor the table of field f is built with: the field's boost when it is +-2^k with a small exponent (exact, see
// ScoreParams::tab_scale), else 1.0.  PB_FOLD_BOOST=0 keeps the multiplication in the loop (A/B, tests).
double boost_fold(double boost) {
  const char* e = std::getenv("PB_FOLD_BOOST");
  const bool off = e && !std::strcmp(e, "0");
  if (off || !std::isfinite(boost) || boost == 0.0 || boost == 1.0) return 1.0;
  int ex = 0;
  const double m = std::frexp(std::fabs(boost), &ex);
  return (m == 0.5 && ex >= -30 && ex <= 31) ? boost : 1.0;
}
bool batch_table_current(const pb_batch* b) {
  if (!(b->tab_valid && b->tab_k1 == b->k1 && b->tab_b == b->b && b->tab_epoch == b->ix->live_epoch)) return false;
  for (uint32_t f = 0; f < b->ix->F; ++f) if (b->tab_scale[f] != boost_fold(b->boost[f])) return false;
  return true;
}

int batch_build_table(pb_batch* b) {
  pb_index* ix = b->ix;
  for (uint32_t f = 0; f < 4; ++f) b->tab_scale[f] = f < ix->F ? boost_fold(b->boost[f]) : 1.0;
  // (tf, fl) -> saturated tf, per field, tf-major with row stride flc.  In the narrow layout the row
  // stride is 1 << fl_bits so that a posting code indexes the table directly; the table may then only
  // be cut along tf.
  uint32_t tfc[4], flc[4];
  for (uint32_t f = 0; f < ix->F; ++f) {
    tfc[f] = std::min<uint32_t>(ix->max_tf[f] + 1, ix->narrow ? 65536u : 64u);
    flc[f] = ix->narrow ? (1u << ix->fl_bits[f]) : std::min<uint32_t>(ix->max_fl[f] + 1, 1024);
  }
  auto total = [&]() { uint64_t t = 0; for (uint32_t f = 0; f < ix->F; ++f) t += (uint64_t)tfc[f] * flc[f]; return t; };
  while (total() > 8192) {     // 64 KB of shared memory
    uint32_t best = 0;
    for (uint32_t f = 1; f < ix->F; ++f) if ((uint64_t)tfc[f] * flc[f] > (uint64_t)tfc[best] * flc[best]) best = f;
    if (tfc[best] > 4) tfc[best] = (tfc[best] + 1) / 2;
    else if (!ix->narrow && flc[best] > 2) flc[best] = (flc[best] + 1) / 2;
    else if (tfc[best] > 1) tfc[best] = (tfc[best] + 1) / 2;
    else break;
  }
  std::vector<double> h;
  uint32_t off = 0;
  for (uint32_t f = 0; f < 4; ++f) { b->tab_tfcap[f] = 0; b->tab_flcap[f] = 0; b->tab_off[f] = 0; }
  for (uint32_t f = 0; f < ix->F; ++f) {
    b->tab_tfcap[f] = tfc[f]; b->tab_flcap[f] = flc[f]; b->tab_off[f] = off;
    for (uint32_t tf = 0; tf < tfc[f]; ++tf)
      for (uint32_t fl = 0; fl < flc[f]; ++fl) h.push_back((tf == 0 ? 0.0 : bm25_tf_host(b->k1, b->b, ix->avg[f], tf, fl)) * b->tab_scale[f]);
    off += tfc[f] * flc[f];
  }
  b->tab_total = off;
  b->tab_full = true;
  for (uint32_t f = 0; f < ix->F; ++f) if (tfc[f] <= ix->max_tf[f] || flc[f] <= ix->max_fl[f]) b->tab_full = false;
  CU(b->tab.ensure(off + 1));
  CU(cudaMemcpyAsync(b->tab.p, h.data(), off * sizeof(double), cudaMemcpyHostToDevice, b->stream));
  CU(cudaStreamSynchronize(b->stream));   // h goes out of scope
  return PB_OK;
}

int batch_load(pb_batch* b, const pb_query_batch_desc* d, uint64_t full_cap, bool sync_inputs = true) {
  pb_index* ix = b->ix;
  if (!d || (d->n_queries && (!d->query_term_off || !d->term_byte_off))) { pb::set_error("query batch: null argument"); return PB_ERR_INVALID; }
  if (d->scorer != PB_SCORER_BM25 && d->scorer != PB_SCORER_ZERO_TO_ONE) {
    pb::set_error("query batch: scorer %u cannot run on the device (only BM25 and ZeroToOne exist there; there is no CPU fallback)", d->scorer);
    return PB_ERR_UNSUPPORTED;
  }
  if (d->top_k > PB_MAX_TOP_K) { pb::set_error("query batch: top_k %u > %u", d->top_k, PB_MAX_TOP_K); return PB_ERR_UNSUPPORTED; }
  if (d->n_fields_boost != ix->F || !d->fields_boost) { pb::set_error("query batch: fields_boost must have %u entries", ix->F); return PB_ERR_INVALID; }
  for (uint32_t f = 0; f < ix->F; ++f)
    if (!std::isfinite(d->fields_boost[f])) { pb::set_error("query batch: fields_boost[%u] is not finite", f); return PB_ERR_INVALID; }
  if (d->n_queries >= 0xFFFFFFF0ull) { pb::set_error("query batch: too many queries"); return PB_ERR_UNSUPPORTED; }
  const uint64_t Q = d->n_queries;
  const uint64_t NT = Q ? d->query_term_off[Q] : 0;
  if (Q && d->query_term_off[0] != 0) { pb::set_error("query batch: query_term_off[0] != 0"); return PB_ERR_INVALID; }
  for (uint64_t q = 0; q < Q; ++q) {
    if (d->query_term_off[q + 1] < d->query_term_off[q]) { pb::set_error("query batch: query_term_off not monotone"); return PB_ERR_INVALID; }
    if (d->query_term_off[q + 1] - d->query_term_off[q] > PB_MAX_QUERY_TERMS) { pb::set_error("query %llu has more than %u terms", (ull)q, PB_MAX_QUERY_TERMS); return PB_ERR_UNSUPPORTED; }
  }
  if (NT && d->term_byte_off[0] != 0) { pb::set_error("query batch: term_byte_off[0] != 0"); return PB_ERR_INVALID; }
  for (uint64_t t = 0; t < NT; ++t) {
    if (d->term_byte_off[t + 1] < d->term_byte_off[t]) { pb::set_error("query batch: term_byte_off not monotone"); return PB_ERR_INVALID; }
    if (!pb::utf8_valid(d->term_bytes + d->term_byte_off[t], d->term_byte_off[t + 1] - d->term_byte_off[t])) {
      pb::set_error("query term %llu is not valid UTF-8", (ull)t);
      return PB_ERR_INVALID;
    }
  }
  const uint64_t NB = NT ? d->term_byte_off[NT] : 0;
  CU(cudaSetDevice(ix->device));
  b->Q = Q; b->NT = NT; b->scorer = d->scorer; b->k = d->top_k; b->k1 = d->bm25_k1; b->b = d->bm25_b;
  for (uint32_t f = 0; f < 4; ++f) b->boost[f] = f < ix->F ? d->fields_boost[f] : 0.0;
  b->full_cap = full_cap;
  // inputs -> device (async on the batch stream; pinned sources overlap, pageable ones stage)
  CU(b->term_bytes.ensure(NB + 16));
  CU(b->term_byte_off.ensure(NT + 2));
  CU(b->query_term_off.ensure(Q + 2));
  if (NB) CU(cudaMemcpyAsync(b->term_bytes.p, d->term_bytes, NB, cudaMemcpyHostToDevice, b->stream));
  if (NT) CU(cudaMemcpyAsync(b->term_byte_off.p, d->term_byte_off, (NT + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, b->stream));
  else CU(cudaMemsetAsync(b->term_byte_off.p, 0, sizeof(uint64_t), b->stream));
  if (Q) CU(cudaMemcpyAsync(b->query_term_off.p, d->query_term_off, (Q + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, b->stream));
  else CU(cudaMemsetAsync(b->query_term_off.p, 0, sizeof(uint64_t), b->stream));
  // workspace that depends only on Q / NT
  CU(b->qt_lo.ensure(NT + 1)); CU(b->qt_hi.ensure(NT + 1)); CU(b->qt_len.ensure(NT + 1)); CU(b->qt_q.ensure(NT + 1));
  CU(b->qt_gcount.ensure(NT + 2)); CU(b->qt_goff.ensure(NT + 2));
  CU(b->q_isg.ensure(Q + 2)); CU(b->q_gidx.ensure(Q + 2)); CU(b->q_grows.ensure(Q + 2)); CU(b->q_prim.ensure(Q + 2));
  CU(b->q_recbound.ensure(Q + 2)); CU(b->q_recoff.ensure(Q + 2)); CU(b->q_nbins.ensure(Q + 2)); CU(b->q_binoff.ensure(Q + 2));
  CU(b->q_scheme.ensure(Q + 2)); CU(b->q_shift.ensure(Q + 2)); CU(b->q_bmwords.ensure(Q + 2)); CU(b->q_bmoff.ensure(Q + 2)); CU(b->q_gsegoff.ensure(Q + 2)); CU(b->q_gtileoff.ensure(Q + 2));
  CU(b->s_tiles.ensure(Q + 2)); CU(b->s_tile_off.ensure(Q + 2));
  CU(b->seg_s.ensure(Q + 1));
  RC(batch_layout_results(b));
  CU(b->part_head.ensure(Q + 1));
  CU(b->scal.ensure(pb_batch::SCAL_WORDS));
  b->counters.p = reinterpret_cast<uint32_t*>(b->scal.p);
  b->stats.p = b->scal.p + 4; b->u_counter.p = b->stats.p + 3 * ST_COUNT; b->full_count.p = b->u_counter.p + 8;
  b->xcount.p = b->full_count.p + 2;
  CU(b->uq.ensure(Q + 1)); CU(b->q_isu.ensure(Q + 2)); CU(b->q_uidx.ensure(Q + 2)); CU(b->u_list.ensure(Q + 1));
  CU(b->q_isu2.ensure(Q + 2)); CU(b->q_uidx2.ensure(Q + 2)); CU(b->u_list2.ensure(Q + 1));
  if (full_cap) {
    CU(b->full_q.ensure(full_cap)); CU(b->full_doc.ensure(full_cap)); CU(b->full_score.ensure(full_cap));
  }
  if (b->scorer == PB_SCORER_BM25 && !batch_table_current(b)) {
    RC(batch_build_table(b));             // (k1, b, avg) unchanged since the last batch staged here: the table stays
    b->tab_valid = true; b->tab_k1 = b->k1; b->tab_b = b->b;
    b->tab_epoch = ix->live_epoch;
  }
  if (sync_inputs) CU(cudaStreamSynchronize(b->stream));   // caller buffers may be reused after return
  b->loaded = true;
  b->ran = false;
  return PB_OK;
}

// Shared-memory plan of the scoring kernel: how many copies of the BM25 table (16 = conflict-free
// LDS.64, see ScoreParams::tab_rep_shift) fit, and the CTA shape that keeps ~24 warps per SM resident.
struct ScorePlan { uint32_t rep_shift; int threads; size_t smem; };
ScorePlan score_plan(const pb_batch* b, uint32_t tab_total) {
  ScorePlan sp{0, 256, 0};
  if (tab_total == 0) return sp;
  const size_t budget = 220u << 10;                 // of the 227 KB a CTA may use
  uint32_t want = 4;                                // 16 copies
  if (const char* e = std::getenv("PB_TAB_REP_SHIFT")) want = (uint32_t)std::min(4, std::max(0, atoi(e)));
  uint32_t sh = want;
  while (sh > 0 && (((size_t)tab_total * 8) << sh) > budget) --sh;
  sp.rep_shift = sh;
  sp.smem = (((size_t)tab_total * 8) << sh) + 128;
  // smallest CTA whose copies still add up to SCORE_MAX_THREADS resident threads per SM within the budget
  sp.threads = SCORE_MAX_THREADS;
  for (int t : {256, SCORE_MAX_THREADS / 2}) {
    if (SCORE_MAX_THREADS % t == 0 && (size_t)(SCORE_MAX_THREADS / t) * (sp.smem + 1024) <= budget) { sp.threads = t; break; }
  }
  (void)b;
  return sp;
}

int launch_score(pb_batch* b, ScoreParams& P, bool gmode, uint64_t tiles) {
  const FieldOps* ops = field_ops(b->ix->F);
  const int sc = (int)b->scorer;
  const bool nw = b->ix->narrow;
  const ScorePlan sp = score_plan(b, P.tab_total);
  P.tab_rep_shift = sp.rep_shift;
  P.tab_stride = 8u << sp.rep_shift;
  for (int x = 0; x < 4; ++x) P.tab_boff[x] = P.tab_off[x] * P.tab_stride;
  int per_sm = 1;
  {
    // cudaFuncSetAttribute + the occupancy query cost several microseconds per launch: remembered per shape
    const uint64_t key = ((uint64_t)sp.smem << 16) | ((uint64_t)sp.threads << 4) | (sc ? 4u : 0u) | (gmode ? 2u : 0u) | (nw ? 1u : 0u);
    std::lock_guard<std::mutex> lk(b->ix->occ_mu);
    auto it = b->ix->occ_cache.find(key);
    if (it == b->ix->occ_cache.end()) {
      CU(ops->score_occupancy(sc, gmode, nw, &per_sm, sp.threads, sp.smem));
      b->ix->occ_cache[key] = per_sm;
    } else {
      per_sm = it->second;
    }
  }
  if (per_sm < 1) { pb::set_error("scoring kernel does not fit an SM (%d threads, %zu B shared)", sp.threads, sp.smem); return PB_ERR_CUDA; }
  // one warp = one contiguous span of tiles; never more warps than there is work for
  const uint64_t warps_per_cta = (uint64_t)sp.threads / 32;
  uint64_t max_grid = (uint64_t)b->ix->sm_count * per_sm;
  uint64_t want = (tiles + warps_per_cta * 2 - 1) / (warps_per_cta * 2);
  int grid = (int)std::max<uint64_t>(1, std::min(max_grid, want));
  CU(ops->score_launch(sc, gmode, nw, &P, grid, sp.threads, sp.smem, b->stream));
  return PB_OK;
}

int launch_mark(pb_batch* b, const ScoreParams& P, int grid, int clear) {
  CU(field_ops(b->ix->F)->mark_launch(&P, clear, grid, b->stream));
  return PB_OK;
}

int launch_fold(pb_batch* b, const FoldParams& FP) {
  uint64_t want = ((uint64_t)FP.n + CTA_THREADS - 1) / CTA_THREADS;
  int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)b->ix->sm_count * 4, want));
  CU(field_ops(b->ix->F)->fold_launch((int)b->scorer, &FP, grid, b->stream));
  return PB_OK;
}

int launch_binfold(pb_batch* b, const ScoreParams& P, uint64_t n_bins) {
  uint64_t want = (n_bins + WARPS_PER_CTA * 4 - 1) / (WARPS_PER_CTA * 4);
  int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)b->ix->sm_count * 6, want));
  CU(field_ops(b->ix->F)->binfold_launch((int)b->scorer, &P, grid, b->stream));
  return PB_OK;
}

// Side-path capacity knobs (bytes of HBM the workspace may take).
constexpr uint64_t REC_CAP_DEFAULT = 128ull << 20;      // records per round (x32 B with sort buffers)
constexpr uint64_t REC_CAP_MAX = 1ull << 31;
constexpr uint64_t BIN_CAP = 96ull << 20;               // doc-range bins per round (12 B each)
constexpr uint64_t BITMAP_POOL_BYTES = 6ull << 30;    // per-query doc bitmaps of one round

int fetch_prefix_arrays(pb_batch* b) {
  if (b->h_full) return PB_OK;
  const uint64_t Q = b->Q;
  cudaStream_t st = b->stream;
  CU(cudaMemcpyAsync(b->h_bmoff.data(), b->q_bmoff.p, (Q + 1) * sizeof(ull), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(b->h_recoff.data(), b->q_recoff.p, (Q + 1) * sizeof(ull), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(b->h_binoff.data(), b->q_binoff.p, (Q + 1) * sizeof(ull), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(b->h_gidx.data(), b->q_gidx.p, (Q + 1) * sizeof(ull), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(b->h_gsegoff.data(), b->q_gsegoff.p, (Q + 1) * sizeof(ull), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(b->h_gtileoff.data(), b->q_gtileoff.p, (Q + 1) * sizeof(ull), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  b->h_full = true;
  return PB_OK;
}

// Stage 1 of a run: every kernel of the batch up to the merged top-k, enqueued on the batch's stream
// (with the host round trips the planning needs).  ev[0] .. ev[5] bracket its phases.
int batch_compute(pb_batch* b) {
  pb_index* ix = b->ix;
  if (!b->loaded) { pb::set_error("pb_batch_run: batch not loaded"); return PB_ERR_INVALID; }
  b->ran = false; b->gathered = false; b->gather_pending = false;
  b->rev_used = 0;
  for (auto& v : b->rev_span) v.clear();
  if (b->scorer == PB_SCORER_BM25 && !batch_table_current(b)) {     // pb_index_set_live_state since staging: new avg
    RC(batch_build_table(b));
    b->tab_epoch = ix->live_epoch;
    b->tab_valid = true; b->tab_k1 = b->k1; b->tab_b = b->b;
  }
  CU(cudaSetDevice(ix->device));
  cudaStream_t st = b->stream;
  const uint64_t Q = b->Q, NT = b->NT;
  const uint32_t k = b->k;
  pb_batch_stats& S = b->st;
  std::memset(&S, 0, sizeof(S));
  S.n_queries = Q; S.n_query_terms = NT;
  uint32_t launches = 0;
  if (ix->n_rows_padded >= 0xFFFFFFFFull) { pb::set_error("index has more than 2^32 posting rows"); return PB_ERR_UNSUPPORTED; }

  CU(cudaEventRecord(b->ev[0], st));
  CU(cudaMemsetAsync(b->res.p, 0, b->block_bytes, st));      // counts, digests, top-k: one packed block
  CU(cudaMemsetAsync(b->part_head.p, 0xFF, (Q + 1) * sizeof(uint32_t), st));
  CU(cudaMemsetAsync(b->scal.p, 0, pb_batch::SCAL_WORDS * sizeof(ull), st));   // counters, stats, work-item counters
  if (Q == 0) {
    for (int i = 1; i <= 5; ++i) CU(cudaEventRecord(b->ev[i], st));
    b->launches = 0;
    return PB_OK;
  }

  IndexView view = ix->view();
  // ---- trie descent + prefix expansion ---------------------------------------------------
  if (NT) {
    descend_kernel<<<(unsigned)((NT + 255) / 256), 256, 0, st>>>(view, b->term_bytes.p, b->term_byte_off.p, NT,
                                                                 b->qt_lo.p, b->qt_hi.p, b->qt_len.p);
    CU(cudaGetLastError());
    ++launches;
  }
  CU(cudaEventRecord(b->ev[1], st));
  // ---- plan ------------------------------------------------------------------------------
  UPlan up;
  std::memset(&up, 0, sizeof(up));
  up.enabled = (b->scorer == PB_SCORER_ZERO_TO_ONE && ix->u_ok) ? 1u : 0u;
  {
    uint64_t div = 16;
    if (const char* e = std::getenv("PB_UNION_MIN_DIV")) div = (uint64_t)std::max<long long>(0, atoll(e));
    up.min_rows = div ? ix->n_docs / div : 0;
  }
  up.uq = b->uq.p; up.q_isu = b->q_isu.p; up.q_isu2 = b->q_isu2.p;
  up.max_term_bytes = ix->max_term_bytes;
  for (uint32_t f = 0; f < ix->F; ++f) up.max_tf = std::max(up.max_tf, ix->max_tf[f]);
  {
    const size_t sm_gen = ix->u_warp ? union_warp_smem_bytes((int)ix->F, ix->u_wbits, true) : union_smem_bytes((int)ix->F, ix->u_wbits, true);
    const size_t sm_fast = ix->u_warp ? union_warp_smem_bytes((int)ix->F, ix->u_wbits, false) : union_smem_bytes((int)ix->F, ix->u_wbits, false);
    up.allow_gen = sm_gen <= (222u << 10) ? 1u : 0u;
    if (sm_fast > (222u << 10)) up.enabled = 0;
  }
  up.liverowcnt_prefix = ix->liverowcnt_prefix.p; up.dflive_prefix = ix->dflive_prefix.p;
  up.stats = b->stats.p + 2 * ST_COUNT;
  plan_query_kernel<<<(unsigned)((Q + 255) / 256), 256, 0, st>>>(view, Q, b->query_term_off.p, b->qt_lo.p, b->qt_hi.p,
                                                                 b->qt_len.p, b->seg_s.p, b->s_tiles.p, b->qt_gcount.p,
                                                                 b->qt_q.p, b->q_isg.p, b->q_grows.p, b->stats.p, up, NT);
  CU(cudaGetLastError());
  ++launches;
  // the exclusive scans below run over n + 1 entries: give the extra input entry a defined value
  RC(scan_ull(b, b->s_tiles.p, b->s_tile_off.p, Q + 1));
  RC(scan_ull(b, b->qt_gcount.p, b->qt_goff.p, NT + 1));
  launches += 2;
  ull h_nu = 0, h_nu2 = 0;            // class-U queries: without / with overlapping term ranges
  if (up.enabled) {
    RC(scan_ull(b, b->q_isu.p, b->q_uidx.p, Q + 1));
    RC(scan_ull(b, b->q_isu2.p, b->q_uidx2.p, Q + 1));
    launches += 2;
    CU(cudaMemcpyAsync(&h_nu, b->q_uidx.p + Q, sizeof(ull), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&h_nu2, b->q_uidx2.p + Q, sizeof(ull), cudaMemcpyDeviceToHost, st));
  }
  ull h_tot[2] = {0, 0};
  CU(cudaMemcpyAsync(&h_tot[0], b->s_tile_off.p + Q, sizeof(ull), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(&h_tot[1], b->qt_goff.p + NT, sizeof(ull), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  const ull s_tiles_total = h_tot[0], n_gsegs = h_tot[1];
  if (n_gsegs >= 0xFFFFFFF0ull) { pb::set_error("batch expands to more than 2^32 posting lists"); return PB_ERR_UNSUPPORTED; }
  S.n_segments = n_gsegs;   // class-S segments are added below from the stats

  struct Round { uint64_t qa, qb, sa, sb, ta, tb, slots, recs, words; };
  std::vector<Round> rounds;
  const uint32_t doc_bits = bits_for(std::max<uint64_t>(ix->n_docs, 2));
  const uint32_t bitmap_sum_words = (uint32_t)(ix->n_docs / 32768 + 1);
  const uint32_t bitmap_doc_words = (uint32_t)((ix->n_docs + 31) / 32 + 1);
  const uint32_t bitmap_words = bitmap_sum_words + 2 * bitmap_doc_words;
  uint64_t rec_cap = 0, pool_words = 0;
  if (n_gsegs) {
    CU(b->seg_g.ensure(n_gsegs + 1));
    CU(b->g_tiles.ensure(n_gsegs + 2));
    CU(b->g_tile_off.ensure(n_gsegs + 2));
    CU(b->g_mtiles.ensure(n_gsegs + 2)); CU(b->g_mtile_off.ensure(n_gsegs + 2));
    CU(cudaMemsetAsync(b->q_prim.p, 0, (Q + 2) * sizeof(ull), st));      // gfill elects the primary list with atomicMax
    gfill_kernel<<<ix->sm_count * 8, 256, 0, st>>>(view, NT, b->query_term_off.p, b->qt_lo.p, b->qt_hi.p, b->qt_len.p,
                                                   b->qt_q.p, b->qt_gcount.p, b->qt_goff.p, b->seg_g.p, b->g_tiles.p,
                                                   b->q_prim.p, b->stats.p + ST_COUNT);
    CU(cudaGetLastError());
    CU(cudaMemsetAsync(b->q_bmwords.p, 0, (Q + 2) * sizeof(ull), st));
    CU(cudaMemsetAsync(b->xcount.p + 2, 0, 2 * sizeof(ull), st));      // per-query maxima: records, mask words
    gprimary_kernel<<<(unsigned)((Q + 1 + 255) / 256), 256, 0, st>>>(Q, doc_bits, b->q_isg.p, b->q_grows.p, b->q_prim.p,
                                                                     b->seg_g.p, b->q_recbound.p, b->q_nbins.p, b->q_scheme.p,
                                                                     b->q_shift.p, b->query_term_off.p, b->qt_goff.p,
                                                                     b->q_gsegoff.p, b->q_bmwords.p, bitmap_words, b->xcount.p + 2);
    CU(cudaGetLastError());
    CU(cudaMemsetAsync(b->g_tiles.p + n_gsegs, 0, sizeof(ull), st));
    CU(cudaMemsetAsync(b->q_recbound.p + Q, 0, sizeof(ull), st));
    CU(cudaMemsetAsync(b->q_nbins.p + Q, 0, sizeof(ull), st));
    CU(cudaMemsetAsync(b->q_isg.p + Q, 0, sizeof(ull), st));
    RC(scan_ull(b, b->g_tiles.p, b->g_tile_off.p, n_gsegs + 1));
    RC(scan_ull(b, b->q_recbound.p, b->q_recoff.p, Q + 1));
    RC(scan_ull(b, b->q_nbins.p, b->q_binoff.p, Q + 1));
    RC(scan_ull(b, b->q_isg.p, b->q_gidx.p, Q + 1));
    RC(scan_ull(b, b->q_bmwords.p, b->q_bmoff.p, Q + 1));
    gather_tileoff_kernel<<<(unsigned)((Q + 1 + 255) / 256), 256, 0, st>>>(Q + 1, b->q_gsegoff.p, b->g_tile_off.p, b->q_gtileoff.p);
    CU(cudaGetLastError());
    launches += 8;
    // Round planning needs the per-query prefix arrays on the host only when the side path does not
    // fit ONE round; the totals and the per-query maxima (8 scalars) decide that.
    for (auto* v : {&b->h_binoff, &b->h_bmoff, &b->h_recoff, &b->h_gidx, &b->h_gsegoff, &b->h_gtileoff}) {
      if (v->size() != Q + 1) v->assign(Q + 1, 0);      // entry [0] of an exclusive prefix is always 0
    }
    ull h_max[2] = {0, 0};
    CU(cudaMemcpyAsync(&b->h_bmoff[Q], b->q_bmoff.p + Q, sizeof(ull), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&b->h_recoff[Q], b->q_recoff.p + Q, sizeof(ull), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&b->h_binoff[Q], b->q_binoff.p + Q, sizeof(ull), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&b->h_gidx[Q], b->q_gidx.p + Q, sizeof(ull), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&b->h_gsegoff[Q], b->q_gsegoff.p + Q, sizeof(ull), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&b->h_gtileoff[Q], b->q_gtileoff.p + Q, sizeof(ull), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_max, b->xcount.p + 2, 2 * sizeof(ull), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    pool_words = std::max<uint64_t>(BITMAP_POOL_BYTES / 4, h_max[1]);       // one query always fits
    rec_cap = std::max<uint64_t>(REC_CAP_DEFAULT, h_max[0]);
    if (rec_cap > REC_CAP_MAX) { pb::set_error("a single query needs %llu side-path records (> %llu)", (ull)rec_cap, (ull)REC_CAP_MAX); return PB_ERR_UNSUPPORTED; }
    b->h_full = false;
    if (b->h_recoff[Q] > rec_cap || b->h_bmoff[Q] > pool_words || b->h_binoff[Q] > BIN_CAP) RC(fetch_prefix_arrays(b));
    uint64_t qa = 0;
    if (!b->h_full) {               // everything fits one round: only entries [0] and [Q] of the arrays are needed
      Round r{0, Q, 0, b->h_gsegoff[Q], 0, b->h_gtileoff[Q], b->h_gidx[Q], b->h_recoff[Q], b->h_bmoff[Q]};
      if (r.sb > r.sa) rounds.push_back(r);
      qa = Q;
    }
    while (qa < Q) {
      // largest qb with recoff[qb]-recoff[qa] <= rec_cap and bmoff[qb]-bmoff[qa] <= pool_words
      uint64_t qb1 = std::upper_bound(b->h_recoff.begin() + qa, b->h_recoff.end(), b->h_recoff[qa] + rec_cap) - b->h_recoff.begin() - 1;
      uint64_t qb2 = std::upper_bound(b->h_bmoff.begin() + qa, b->h_bmoff.end(), b->h_bmoff[qa] + pool_words) - b->h_bmoff.begin() - 1;
      uint64_t qb3 = std::upper_bound(b->h_binoff.begin() + qa, b->h_binoff.end(), b->h_binoff[qa] + BIN_CAP) - b->h_binoff.begin() - 1;
      uint64_t qb = std::max<uint64_t>(qa + 1, std::min(qb1, std::min(qb2, qb3)));
      Round r{qa, qb, b->h_gsegoff[qa], b->h_gsegoff[qb], b->h_gtileoff[qa], b->h_gtileoff[qb],
              b->h_gidx[qb] - b->h_gidx[qa], b->h_recoff[qb] - b->h_recoff[qa], b->h_bmoff[qb] - b->h_bmoff[qa]};
      if (r.sb > r.sa) rounds.push_back(r);
      qa = qb;
    }
  }
  S.side_rounds = (uint32_t)rounds.size();

  // partial top-k lists: <= 2 per warp per launch + 2 per class-G query
  const uint64_t max_warps = (uint64_t)ix->sm_count * 8 * WARPS_PER_CTA;
  const uint64_t n_gq = n_gsegs ? b->h_gidx[Q] : 0;
  const uint64_t u_spi = ix->u_warp ? (U_ITEM_DOCS >> ix->u_wbits) : (uint64_t)U_CHUNK;      // shards per work item
  const uint64_t u_chunks = ix->u_ok ? (ix->u_shards + u_spi - 1) / u_spi : 0;
  const uint64_t u_items = (h_nu + h_nu2) * u_chunks;
  uint64_t part_cap = 2 * max_warps * (1 + 2 * rounds.size()) + 2 * n_gq + 64 + u_items;
  if (part_cap >= 0xFFFFFFF0ull) { pb::set_error("too many partial lists"); return PB_ERR_UNSUPPORTED; }
  const uint32_t kk = std::max<uint32_t>(k, 1);
  CU(b->part_next.ensure(part_cap)); CU(b->part_n.ensure(part_cap));
  CU(b->part_doc.ensure(part_cap * kk));
  CU(b->part_score.ensure(part_cap * kk));
  uint64_t part_bound = 2 * max_warps + u_items;      // host-side upper bound of the device's partial-list counter

  ScoreParams P;
  std::memset(&P, 0, sizeof(P));
  P.ix = view;
  P.out.n_results = b->n_results; P.out.doc_digest = b->doc_digest; P.out.score_digest = b->score_digest;
  P.out.topk_n = b->topk_n; P.out.topk_doc = b->topk_doc; P.out.topk_score = b->topk_score; P.out.k = k;
  P.out.part_head = b->part_head.p; P.out.part_next = b->part_next.p; P.out.part_n = b->part_n.p;
  P.out.part_doc = b->part_doc.p; P.out.part_score = b->part_score.p; P.out.part_count = b->counters.p + 0;
  P.out.part_cap = (uint32_t)part_cap;
  if (b->full_cap) { P.out.full_q = b->full_q.p; P.out.full_doc = b->full_doc.p; P.out.full_score = b->full_score.p; }
  P.out.full_count = b->full_count.p; P.out.full_cap = b->full_cap;
  P.out.error_flag = b->counters.p + 2;
  P.out.results_total = b->full_count.p + 1;
  P.query_term_off = b->query_term_off.p;
  P.k1 = b->k1; P.b = b->b; P.one_minus_b = 1.0 - b->b; P.k1_plus_1 = b->k1 + 1.0;
  for (int f = 0; f < 4; ++f) { P.tab_scale[f] = b->scorer == PB_SCORER_BM25 ? b->tab_scale[f] : 1.0; P.boost[f] = P.tab_scale[f] != 1.0 ? 1.0 : b->boost[f]; P.avg[f] = ix->avg[f]; P.tab_tfcap[f] = b->tab_tfcap[f]; P.tab_flcap[f] = b->tab_flcap[f]; P.tab_off[f] = b->tab_off[f]; }
  P.tab = b->tab.p; P.tab_total = b->scorer == 0 ? b->tab_total : 0;
  P.tab_full = b->tab_full ? 1u : 0u;
  // tiles PB_L2_AHEAD ahead are pulled into L2 (measured on cfg 1, whose image is only 207 MB: 25.1 ms with the prefetch,
  // 28.6 ms without — it pays even when 90 % of the sectors hit L2); PB_L2_PREFETCH=0 switches it off for experiments
  P.l2_prefetch = 1u;
  if (const char* e = std::getenv("PB_L2_PREFETCH")) P.l2_prefetch = atoi(e) ? 1u : 0u;
  P.boosts_all_one = 1u;
  for (uint32_t f = 0; f < ix->F; ++f) if (P.boost[f] != 1.0) P.boosts_all_one = 0u;      // folded boosts count as 1.0
  P.doc_bits = doc_bits; P.bitmap_words = bitmap_words; P.bitmap_sum_words = bitmap_sum_words; P.bitmap_doc_words = bitmap_doc_words;
  P.xcount = b->xcount.p; P.xtiles = b->xcount.p + 1;
  P.q_bmoff = b->q_bmoff.p; P.q_prim = b->q_prim.p;
  P.q_binoff = b->q_binoff.p; P.q_shift = b->q_shift.p; P.q_gsegoff = b->q_gsegoff.p;
  P.rec_count = b->counters.p + 1;
  CU(cudaEventRecord(b->ev[2], st));

  // ---- class S: every single-list query in ONE launch of the scoring kernel --------------
  if (s_tiles_total) {
    P.segs = b->seg_s.p; P.tile_off = (const uint64_t*)b->s_tile_off.p;
    P.seg_begin = 0; P.seg_end = (uint32_t)Q; P.tile_begin = 0; P.tile_end = s_tiles_total;
    P.stats = b->stats.p;
    RC(launch_score(b, P, false, s_tiles_total));
    ++launches;
    S.score_launches = 1;
  }
  CU(cudaEventRecord(b->ev[3], st));

  // ---- class U: union-heavy ZeroToOne queries, one dense pass (union_kernels.cuh) -----------
  for (int gen = 0; gen < 2; ++gen) {
    const ull nu = gen ? h_nu2 : h_nu;
    if (!nu) continue;
    const ull* isu = gen ? b->q_isu2.p : b->q_isu.p;
    const ull* uidx = gen ? b->q_uidx2.p : b->q_uidx.p;
    uint32_t* list = gen ? b->u_list2.p : b->u_list.p;
    ucompact_kernel<<<(unsigned)((Q + 255) / 256), 256, 0, st>>>(Q, isu, uidx, list);
    CU(cudaGetLastError());
    UnionParams UP;
    std::memset(&UP, 0, sizeof(UP));
    UP.ix = view; UP.uv = ix->uview(); UP.out = P.out;
    UP.uq = b->uq.p; UP.u_list = list; UP.n_u = (uint32_t)nu; UP.n_chunks = (uint32_t)u_chunks;
    UP.item_counter = b->u_counter.p + gen;
    UP.prof = b->u_counter.p + 2;
    const size_t smem = ix->u_warp ? union_warp_smem_bytes((int)ix->F, ix->u_wbits, gen != 0) : union_smem_bytes((int)ix->F, ix->u_wbits, gen != 0);
    int per_sm = 0;
    if (ix->u_warp) CU(field_ops(ix->F)->union_warp_occupancy(gen != 0, &per_sm, smem));
    else CU(field_ops(ix->F)->union_occupancy(gen != 0, &per_sm, smem));
    if (per_sm < 1) { pb::set_error("union kernel does not fit an SM (%zu B shared)", smem); return PB_ERR_CUDA; }
    const int grid = (int)std::min<uint64_t>((uint64_t)ix->sm_count * per_sm, nu * u_chunks);
    const int e0 = b->rev_begin();
    if (ix->u_warp) CU(field_ops(ix->F)->union_warp_launch(gen != 0, &UP, grid, smem, st));
    else CU(field_ops(ix->F)->union_launch(gen != 0, &UP, grid, smem, st));
    b->rev_end(3, e0);
    launches += 2;
    S.union_queries += nu;
  }

  // ---- class G: rounds of the side path --------------------------------------------------
  if (!rounds.empty()) {
    uint64_t bm_words_total = 0;
    for (auto& r : rounds) bm_words_total = std::max(bm_words_total, r.words);
    CU(b->bitmap.ensure(bm_words_total + 8));
    if (b->bitmap_zeroed < b->bitmap.cap) {      // freshly (re)allocated: zero once; rounds clean up after themselves
      CU(cudaMemsetAsync(b->bitmap.p, 0, b->bitmap.cap * sizeof(uint32_t), st));
      b->bitmap_zeroed = b->bitmap.cap;
    }
    CU(b->rec.ensure(rec_cap));
    P.segs = b->seg_g.p; P.tile_off = (const uint64_t*)b->g_tile_off.p;
    P.bitmap = b->bitmap.p;
    P.rec = b->rec.p;
    P.stats = b->stats.p + ST_COUNT;
    uint64_t legacy_cap = 0;
    // Rounds are planned with an ESTIMATE for exact-scheme queries; the marking pass counts the
    // real capacity of every bin, and a round that would overflow the record buffer is split (its
    // marks are cleared first) before anything has been scored.
    auto make_round = [&](uint64_t qa, uint64_t qb) {
      return Round{qa, qb, b->h_gsegoff[qa], b->h_gsegoff[qb], b->h_gtileoff[qa], b->h_gtileoff[qb],
                   b->h_gidx[qb] - b->h_gidx[qa], b->h_recoff[qb] - b->h_recoff[qa], b->h_bmoff[qb] - b->h_bmoff[qa]};
    };
    std::vector<Round> work(rounds.rbegin(), rounds.rend());
    uint32_t n_rounds = 0;
    while (!work.empty()) {
      Round r = work.back();
      work.pop_back();
      if (r.sb == r.sa) continue;
      P.seg_begin = (uint32_t)r.sa; P.seg_end = (uint32_t)r.sb; P.tile_begin = r.ta; P.tile_end = r.tb;
      const uint64_t nseg = r.sb - r.sa, tiles = r.tb - r.ta;
      const uint64_t n_bins = b->h_binoff[r.qb] - b->h_binoff[r.qa];
      if (n_bins >= 0xFFFFFFF0ull) { pb::set_error("too many side-path bins in one round"); return PB_ERR_UNSUPPORTED; }
      CU(b->bin_count.ensure(n_bins + 2)); CU(b->bin_off.ensure(n_bins + 2)); CU(b->bin_cursor.ensure(n_bins + 2));
      P.round_bin0 = b->h_binoff[r.qa]; P.n_bins = (uint32_t)n_bins;
      P.round_bm0 = b->h_bmoff[r.qa];
      P.bin_count = b->bin_count.p; P.bin_off = b->bin_off.p; P.bin_cursor = b->bin_cursor.p;
      CU(cudaMemsetAsync(b->bin_count.p, 0, (n_bins + 2) * sizeof(uint32_t), st));
      CU(cudaMemsetAsync(b->bin_cursor.p, 0, (n_bins + 2) * sizeof(uint32_t), st));
      gslot_kernel<<<(unsigned)((nseg + 1 + 255) / 256), 256, 0, st>>>(b->seg_g.p, r.sa, r.sb, b->q_gidx.p, b->q_scheme.p, (uint32_t)r.qa,
                                                                       b->g_tiles.p, b->g_mtiles.p);
      CU(cudaGetLastError());
      RC(scan_ull(b, b->g_mtiles.p + r.sa, b->g_mtile_off.p + r.sa, nseg + 1));
      launches += 2;
      ScoreParams PM = P;                      // the marking pass walks its own tile space
      PM.tile_off = (const uint64_t*)b->g_mtile_off.p;
      int mgrid = ix->sm_count * 8;            // the size of the marking tile space is only known on the device
      CU(cudaMemsetAsync(b->xcount.p, 0, 2 * sizeof(ull), st));
      { const int e0 = b->rev_begin(); RC(launch_mark(b, PM, mgrid, 0)); b->rev_end(0, e0); }
      launches += 2;
      ull h_x[2] = {0, 0};
      CU(cudaMemcpyAsync(h_x, b->xcount.p, 2 * sizeof(ull), cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      const uint64_t need = h_x[0];             // = sum of the bin capacities
      const uint64_t exact_tiles = h_x[1];      // tiles of exact-scheme lists (their doc bitmaps need clearing)
      auto clear_marks = [&](bool scored) -> int {
        // Row masks of the primary scheme are cleared by the scoring pass itself.  The doc bitmaps of
        // the exact scheme are undone by re-walking the marked rows, or the round's whole pool range
        // is wiped when that is less traffic (and always when nothing was scored).
        if (scored && exact_tiles == 0) return PB_OK;
        if (!scored || exact_tiles * TILE_ROWS * 4 > r.words * 4) {
          CU(cudaMemsetAsync(b->bitmap.p, 0, (size_t)(r.words + 4) * sizeof(uint32_t), st));
        } else {
          RC(launch_mark(b, PM, mgrid, 1));
          ++launches;
        }
        return PB_OK;
      };
      if (need > rec_cap) {
        RC(clear_marks(false));                 // nothing was scored yet
        if (r.qb - r.qa > 1) {
          RC(fetch_prefix_arrays(b));           // splitting needs the per-query prefixes
          uint64_t mid = r.qa + (r.qb - r.qa) / 2;
          work.push_back(make_round(mid, r.qb));
          work.push_back(make_round(r.qa, mid));
          continue;
        }
        if (need > REC_CAP_MAX) { pb::set_error("a single query needs %llu side-path records (> %llu)", (ull)need, (ull)REC_CAP_MAX); return PB_ERR_UNSUPPORTED; }
        rec_cap = need + 1024;
        CU(b->rec.ensure(rec_cap));
        P.rec = b->rec.p;
        work.push_back(r);
        continue;
      }
      ++n_rounds;
      // partial top-k lists this round can add: <= 2 per warp per launch (score, bin fold, legacy fold) + 2 per query
      part_bound += 6 * max_warps + 2 * r.slots;
      if (part_bound > part_cap) {           // rounds were split: grow, keeping the lists written so far
        const uint64_t ncap = part_bound * 2;
        if (ncap >= 0xFFFFFFF0ull) { pb::set_error("too many partial lists"); return PB_ERR_UNSUPPORTED; }
        CU(b->part_next.ensure_keep(ncap, part_cap, st)); CU(b->part_n.ensure_keep(ncap, part_cap, st));
        CU(b->part_doc.ensure_keep(ncap * kk, part_cap * kk, st));
        CU(b->part_score.ensure_keep(ncap * kk, part_cap * kk, st));
        part_cap = ncap;
        P.out.part_next = b->part_next.p; P.out.part_n = b->part_n.p; P.out.part_doc = b->part_doc.p;
        P.out.part_score = b->part_score.p; P.out.part_cap = (uint32_t)part_cap;
      }
      if (need) {
        RC(scan_u32(b, b->bin_count.p, b->bin_off.p, n_bins + 1));
        ++launches;
      } else {
        CU(cudaMemsetAsync(b->bin_off.p, 0, (n_bins + 2) * sizeof(uint32_t), st));
      }
      { const int e0 = b->rev_begin(); RC(launch_score(b, P, true, tiles)); b->rev_end(1, e0); }
      ++launches;
      if (need) {
        // records of bins that overflow a warp window go through the legacy sorted path
        overflow_count_kernel<<<(unsigned)std::min<uint64_t>(1024, (n_bins + 255) / 256), 256, 0, st>>>(b->bin_cursor.p, (uint32_t)n_bins, b->counters.p + 3);
        CU(cudaGetLastError());
        uint32_t h_over = 0;
        CU(cudaMemcpyAsync(&h_over, b->counters.p + 3, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        CU(cudaMemsetAsync(b->counters.p + 3, 0, sizeof(uint32_t), st));
        if (std::getenv("PB_DEBUG"))
          fprintf(stderr, "[pb] round q[%llu,%llu) segs %llu tiles %llu bins %llu need %llu overflow %u exact_tiles %llu words %llu\n",
                  (ull)r.qa, (ull)r.qb, (ull)nseg, (ull)tiles, (ull)n_bins, (ull)need, h_over, (ull)exact_tiles, (ull)r.words);
        if (h_over > legacy_cap) {
          legacy_cap = (uint64_t)h_over + h_over / 4 + 1024;
          CU(b->rec_key.ensure(legacy_cap)); CU(b->rec_val.ensure(legacy_cap));
          CU(b->rec_key2.ensure(legacy_cap)); CU(b->rec_val2.ensure(legacy_cap)); CU(b->rec_flags.ensure(legacy_cap));
        }
        P.rec_key = b->rec_key.p; P.rec_val = b->rec_val.p; P.rec_cap = (uint32_t)legacy_cap;
        const int ef0 = b->rev_begin();
        RC(launch_binfold(b, P, n_bins));
        launches += 2;
        if (h_over) {
          const int end_bit = (int)std::min<uint32_t>(64, doc_bits + bits_for(std::max<uint64_t>(r.slots, 2)));
          size_t bytes = 0;
          CU(cub::DeviceRadixSort::SortPairs(nullptr, bytes, b->rec_key.p, b->rec_key2.p, b->rec_val.p, b->rec_val2.p,
                                             (int64_t)h_over, 0, end_bit, st));
          CU(b->cub_temp.ensure(bytes));
          bytes = b->cub_temp.cap;
          CU(cub::DeviceRadixSort::SortPairs(b->cub_temp.p, bytes, b->rec_key.p, b->rec_key2.p, b->rec_val.p, b->rec_val2.p,
                                             (int64_t)h_over, 0, end_bit, st));
          FoldParams FP;
          FP.S = P; FP.key = b->rec_key2.p; FP.val = b->rec_val2.p; FP.flags = b->rec_flags.p; FP.n = h_over;
          RC(launch_fold(b, FP));
          launches += 2 + (uint32_t)((end_bit + 7) / 8);
          S.legacy_records += h_over;
        }
        b->rev_end(2, ef0);
      }
      RC(clear_marks(true));
      CU(cudaMemsetAsync(b->counters.p + 1, 0, sizeof(uint32_t), st));
    }
    S.side_rounds = n_rounds;
  }
  CU(cudaEventRecord(b->ev[4], st));

  // ---- finalize: merge partial top-k lists -----------------------------------------------
  if (k) {
    uint64_t want = (Q + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)ix->sm_count * 8, want));
    finalize_kernel<<<grid, CTA_THREADS, 0, st>>>(P.out, Q);
    CU(cudaGetLastError());
    ++launches;
  }
  CU(cudaEventRecord(b->ev[5], st));
  b->launches = launches;
  return PB_OK;
}

// Stage 2 (multi-GPU only): ONE ncclAllGather of the packed result block on the batch's stream, right
// behind the last kernel — the next run's memsets are ordered after it on the same stream.
int batch_gather(pb_batch* b) {
  if (!b->comm) { pb::set_error("pb_batch_gather: no communicator attached (pb_batch_set_gather)"); return PB_ERR_INVALID; }
  CU(cudaSetDevice(b->ix->device));
  CU(cudaEventRecord(b->ev[6], b->stream));
  RC(pbg::comm_allgather(b->comm, b->res.p, b->gath.p, b->block_bytes, b->stream));
  CU(cudaEventRecord(b->ev[7], b->stream));
  b->gather_pending = true;
  return PB_OK;
}

// Stage 3: wait for the stream, read the device-side counters and error flags, fill the stats.
int batch_finish(pb_batch* b) {
  pb_index* ix = b->ix;
  cudaStream_t st = b->stream;
  pb_batch_stats& S = b->st;
  CU(cudaSetDevice(ix->device));
  ull h_scal[pb_batch::SCAL_WORDS];
  CU(cudaMemcpyAsync(h_scal, b->scal.p, sizeof(h_scal), cudaMemcpyDeviceToHost, st));      // counters + stats in one copy
  CU(cudaEventRecord(b->ev[8], st));
  CU(cudaStreamSynchronize(st));
  uint32_t h_cnt[4];
  std::memcpy(h_cnt, h_scal, sizeof(h_cnt));
  const ull* h_stats = h_scal + 4;
  const ull* h_full = h_scal + 4 + 3 * ST_COUNT + 8;
  if (h_cnt[2] & 1u) { pb::set_error("internal: partial top-k list overflow"); return PB_ERR_INVALID; }
  if (h_cnt[2] & 2u) { pb::set_error("internal: side-path record buffer overflow"); return PB_ERR_INVALID; }
  if (h_cnt[2] & 8u) { pb::set_error("a query expands to more than 2^27 posting lists"); return PB_ERR_UNSUPPORTED; }
  S.rows_streamed = h_stats[ST_ROWS_STREAMED] + h_stats[ST_COUNT + ST_ROWS_STREAMED] + h_stats[2 * ST_COUNT + ST_ROWS_STREAMED];
  S.rows_scored = h_stats[ST_ROWS_SCORED] + h_stats[ST_COUNT + ST_ROWS_SCORED] + h_stats[2 * ST_COUNT + ST_ROWS_SCORED];
  S.pointer_visits = h_stats[ST_POINTER_VISITS] + h_stats[ST_COUNT + ST_POINTER_VISITS] + h_stats[2 * ST_COUNT + ST_POINTER_VISITS];
  S.rows_streamed_union = h_stats[2 * ST_COUNT + ST_ROWS_STREAMED];
  S.rows_diverted = h_stats[ST_COUNT + ST_ROWS_DIVERTED];
  S.rows_streamed_direct = h_stats[ST_ROWS_STREAMED];
  S.rows_streamed_compact = h_stats[ST_ROWS_COMPACT];
  S.results_emitted = h_full[1];
  S.gpu_launches = b->launches + (b->gather_pending ? 1u : 0u);
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, b->ev[0], b->ev[8])); S.ms_total = ms;
  CU(cudaEventElapsedTime(&ms, b->ev[0], b->ev[1])); S.ms_descend = ms;
  CU(cudaEventElapsedTime(&ms, b->ev[1], b->ev[2])); S.ms_plan = ms;
  CU(cudaEventElapsedTime(&ms, b->ev[2], b->ev[3])); S.ms_score = ms;
  CU(cudaEventElapsedTime(&ms, b->ev[3], b->ev[4])); S.ms_side = ms;
  CU(cudaEventElapsedTime(&ms, b->ev[4], b->ev[5])); S.ms_finalize = ms;
  {
    float* dst[4] = {&S.ms_side_mark, &S.ms_side_score, &S.ms_side_fold, &S.ms_union};
    for (int c = 0; c < 4; ++c) {
      double acc = 0;
      for (auto& pr : b->rev_span[c]) { CU(cudaEventElapsedTime(&ms, b->rev[pr.first], b->rev[pr.second])); acc += ms; }
      *dst[c] = (float)acc;
    }
  }
  S.rows_streamed_side = h_stats[ST_COUNT + ST_ROWS_STREAMED];
  if (std::getenv("PB_UNION_PROF") && S.union_queries) {
    const ull* h_prof = h_scal + 4 + 3 * ST_COUNT + 2;
    fprintf(stderr, "[pb] union kernel cycles (thread 0 of every CTA): setup %llu count %llu score %llu merge-docs %llu item-end %llu, items %llu\n",
            h_prof[0], h_prof[1], h_prof[2], h_prof[5], h_prof[3], h_prof[4]);
  }
  S.ms_gather = 0.f;
  if (b->gather_pending) {
    CU(cudaEventElapsedTime(&ms, b->ev[6], b->ev[7])); S.ms_gather = ms;
    b->gathered = true;
    b->gather_pending = false;
  }
  b->ran = true;
  return PB_OK;
}

int batch_run(pb_batch* b) {
  RC(batch_compute(b));
  if (b->comm) RC(batch_gather(b));
  return batch_finish(b);
}

int batch_new(pb_index* ix, pb_batch** out) {
  CU(cudaSetDevice(ix->device));
  pb_batch* b = new pb_batch();
  b->ix = ix;
  cudaError_t e = cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking);
  for (int i = 0; i < 9 && e == cudaSuccess; ++i) e = cudaEventCreate(&b->ev[i]);
  if (e != cudaSuccess) { delete b; pb::set_error("stream/event creation failed: %s", cudaGetErrorString(e)); return PB_ERR_CUDA; }
  *out = b;
  return PB_OK;
}

int batch_fetch_enqueue(pb_batch* b, pb_query_results* o);

int batch_fetch(pb_batch* b, pb_query_results* o) {
  if (!b->ran) { pb::set_error("pb_batch_fetch: batch has not been run"); return PB_ERR_INVALID; }
  CU(cudaSetDevice(b->ix->device));
  RC(batch_fetch_enqueue(b, o));
  CU(cudaStreamSynchronize(b->stream));
  return PB_OK;
}

// the device-to-host copies of the per-query results, enqueued on the batch's stream (no wait)
int batch_fetch_enqueue(pb_batch* b, pb_query_results* o) {
  const uint64_t Q = b->Q;
  cudaStream_t st = b->stream;
  if (Q) {
    if (o->n_results) CU(cudaMemcpyAsync(o->n_results, b->n_results, Q * sizeof(ull), cudaMemcpyDeviceToHost, st));
    if (o->doc_digest) CU(cudaMemcpyAsync(o->doc_digest, b->doc_digest, Q * sizeof(ull), cudaMemcpyDeviceToHost, st));
    if (o->score_digest) CU(cudaMemcpyAsync(o->score_digest, b->score_digest, Q * sizeof(ull), cudaMemcpyDeviceToHost, st));
    if (o->topk_n) CU(cudaMemcpyAsync(o->topk_n, b->topk_n, Q * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    if (b->k && o->topk_doc) CU(cudaMemcpyAsync(o->topk_doc, b->topk_doc, Q * b->k * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    if (b->k && o->topk_score) CU(cudaMemcpyAsync(o->topk_score, b->topk_score, Q * b->k * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  return PB_OK;
}

// Exclusive side of pb_index::state_mu: announces itself first, so that new one-call queries wait behind it.
struct StateChange {
  pb_index* ix;
  std::unique_lock<std::shared_mutex> wr;
  explicit StateChange(pb_index* i) : ix(i), wr(i->state_mu, std::defer_lock) {
    ix->writers_waiting.fetch_add(1, std::memory_order_acq_rel);
    wr.lock();
    ix->writers_waiting.fetch_sub(1, std::memory_order_acq_rel);
  }
};

// One internal batch, borrowed for the duration of a one-call query (see pb_index::state_mu).  At most PB_QUERY_SLOTS
// (default 4) exist per index; a fifth concurrent caller waits for one to come back.
struct ScratchLease {
  pb_index* ix;
  pb_batch* b = nullptr;
  std::shared_lock<std::shared_mutex> rd;
  explicit ScratchLease(pb_index* i) : ix(i), rd(i->state_mu, std::defer_lock) {
    while (ix->writers_waiting.load(std::memory_order_acquire) > 0) std::this_thread::yield();
    rd.lock();
  }
  ScratchLease(const ScratchLease&) = delete;
  ScratchLease& operator=(const ScratchLease&) = delete;
  int acquire() {
    static const int max_slots = [] { const char* e = std::getenv("PB_QUERY_SLOTS"); return e ? std::min(16, std::max(1, atoi(e))) : 4; }();
    std::unique_lock<std::mutex> lk(ix->pool_mu);
    for (;;) {
      if (!ix->pool_free.empty()) { b = ix->pool_free.back(); ix->pool_free.pop_back(); return PB_OK; }
      if (ix->pool_size < max_slots) {
        ++ix->pool_size;
        lk.unlock();
        int rc = batch_new(ix, &b);
        if (rc != PB_OK) { lk.lock(); --ix->pool_size; b = nullptr; ix->pool_cv.notify_one(); }
        return rc;
      }
      ix->pool_cv.wait(lk);
    }
  }
  ~ScratchLease() {
    if (!b) return;
    std::lock_guard<std::mutex> lk(ix->pool_mu);
    ix->last_st = b->st; ix->has_last = true;
    ix->pool_free.push_back(b);
    ix->pool_cv.notify_one();
  }
};

}  // namespace

// Each segment learns the other's per-term live occurrence counts (matched by the builder's stable term ids) and
// recomputes its idf: the reference has ONE posting list per term, so BM25's document frequency is the sum.
static int segments_exchange_df(pb_index* m) {
  pb_index* d = m->delta;
  if (!d) { m->h_df_extra.clear(); return index_upload_idf(m); }
  uint32_t n_sid = 0;
  for (uint32_t v : m->sid) n_sid = std::max(n_sid, v + 1);
  for (uint32_t v : d->sid) n_sid = std::max(n_sid, v + 1);
  std::vector<uint64_t> by_m(n_sid, 0), by_d(n_sid, 0);
  for (size_t t = 0; t < m->sid.size(); ++t) by_m[m->sid[t]] = m->h_df_live[t];
  for (size_t t = 0; t < d->sid.size(); ++t) by_d[d->sid[t]] = d->h_df_live[t];
  m->h_df_extra.resize(m->sid.size());
  d->h_df_extra.resize(d->sid.size());
  for (size_t t = 0; t < m->sid.size(); ++t) m->h_df_extra[t] = by_d[m->sid[t]];
  for (size_t t = 0; t < d->sid.size(); ++t) d->h_df_extra[t] = by_m[d->sid[t]];
  ++m->live_epoch; ++d->live_epoch;
  CU(cudaSetDevice(m->device));
  RC(index_upload_idf(m));
  return index_upload_idf(d);
}

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

int pb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

static int index_create_impl(const pb_index_image* im, const pb::BuilderLogView* log, int device, pb_index** out);

int pb_index_create(const pb_index_image* im, int device, pb_index** out) {
  if (!im || !out) { pb::set_error("pb_index_create: null argument"); return PB_ERR_INVALID; }
  if (!im->post_blocks) { pb::set_error("pb_index_create: the image has no posting columns (a structure-only flatten goes through pb_index_create_from_builder)"); return PB_ERR_INVALID; }
  return index_create_impl(im, nullptr, device, out);
}

int pb_index_create_from_builder(pb_builder* b, uint64_t from_doc_ordinal, int device, pb_index** out) {
  if (!b || !out) { pb::set_error("pb_index_create_from_builder: null argument"); return PB_ERR_INVALID; }
  PB_TRY({
    pb_index_image im;
    pb::BuilderLogView log;
    RC(pb::builder_flatten_structure(b, from_doc_ordinal, &im, &log));
    return index_create_impl(&im, &log, device, out);
  });
}

int pb_builder_flatten_structure(pb_builder* b, uint64_t from_doc_ordinal, pb_index_image* out) {
  if (!b || !out) { pb::set_error("pb_builder_flatten_structure: null argument"); return PB_ERR_INVALID; }
  PB_TRY({
    pb::BuilderLogView log;
    return pb::builder_flatten_structure(b, from_doc_ordinal, out, &log);
  });
}

// `log` != NULL: the posting columns are flattened ON THE DEVICE from the builder's append log (the image carries none)
static int index_create_impl(const pb_index_image* im, const pb::BuilderLogView* log, int device, pb_index** out) {
  if (im->version != 1 || im->num_fields == 0 || im->num_fields > PB_MAX_FIELDS) { pb::set_error("pb_index_create: bad image header"); return PB_ERR_INVALID; }
  if (im->n_rows_padded % TILE_ROWS != 0 || im->n_rows_padded < im->n_rows + TILE_ROWS) { pb::set_error("pb_index_create: posting columns must be padded to whole 128-row tiles plus one spare tile"); return PB_ERR_INVALID; }
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    pb::set_error("no CUDA device is available (%s); this library has no CPU fallback", ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
    return PB_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= ndev) { pb::set_error("pb_index_create: device %d out of range (0..%d)", device, ndev - 1); return PB_ERR_INVALID; }
  PB_TRY({
    // the kernels index with the image's offsets / ordinals unchecked and the u16 posting codes are sized from
    // max_tf / max_fl: refuse an inconsistent image here (PB_TRUST_IMAGE=1 skips the O(rows) pass).  The builder's own
    // output (log != NULL) is trusted.
    if (!log && !(std::getenv("PB_TRUST_IMAGE") && !std::strcmp(std::getenv("PB_TRUST_IMAGE"), "1"))) RC(pb::validate_image(im));
    CU(cudaSetDevice(device));
    pb_index* ix = new pb_index();
    std::unique_ptr<pb_index> guard(ix);
    ix->device = device; ix->sm_count = g_sm_count(device);
    ix->F = im->num_fields;
    ix->n_nodes = im->n_nodes; ix->n_edges = im->n_edges; ix->n_terms = im->n_terms;
    ix->n_rows = im->n_rows; ix->n_rows_padded = im->n_rows_padded; ix->n_docs = im->n_docs;
    ix->max_term_bytes = im->max_term_bytes;
    for (uint32_t f = 0; f < ix->F; ++f) { ix->max_tf[f] = im->max_tf[f]; ix->max_fl[f] = im->max_fl[f]; }
    CU(upload(ix->node_edge_begin, im->node_edge_begin, im->n_nodes + 1));
    CU(upload(ix->node_term_lo, im->node_term_lo, im->n_nodes));
    CU(upload(ix->node_term_hi, im->node_term_hi, im->n_nodes));
    CU(upload(ix->edge_char, im->edge_char, im->n_edges));
    CU(upload(ix->edge_child, im->edge_child, im->n_edges));
    CU(upload(ix->term_row_begin, im->term_row_begin, im->n_terms + 1));
    CU(upload(ix->term_byte_len, im->term_byte_len, im->n_terms));
    // Posting columns: when, per field, (tf, field length) packs into 16 bits, the device copy keeps one
    // u16 code = tf << fl_bits | fl per field (4 + 2F bytes per row instead of 4 + 8F); the code is
    // also the index into the field's BM25 table.  PB_POSTING_LAYOUT=wide|narrow|auto (default auto).
    {
      // device-side flatten, part 1: sort keys + the exact maxima that decide the layout
      DBuf<uint8_t> d_log;
      DBuf<uint32_t> d_doc_fl, d_ord, d_keys, d_vals, d_keys2, d_vals2, d_max;
      if (log) {
        const uint64_t n = log->n_tuples;
        if (n != im->n_rows) { pb::set_error("pb_index_create_from_builder: the log holds %llu tuples, the image %llu rows", (ull)n, (ull)im->n_rows); return PB_ERR_INVALID; }
        CU(d_log.ensure(std::max<uint64_t>(n, 1) * sizeof(pb::LogTuple)));
        if (n) CU(cudaMemcpy(d_log.p, log->tuples, n * sizeof(pb::LogTuple), cudaMemcpyHostToDevice));
        CU(upload(d_doc_fl, log->doc_fl, log->n_docs * log->F));
        CU(upload(d_ord, log->ord_of, log->n_ord));
        CU(d_keys.ensure(n + 1)); CU(d_vals.ensure(n + 1)); CU(d_keys2.ensure(n + 1)); CU(d_vals2.ensure(n + 1));
        CU(d_max.ensure(8));
        if (n) {
          flat_keys_kernel<<<ix->sm_count * 8, 256>>>(reinterpret_cast<const FlatTuple*>(d_log.p), n, d_ord.p, d_doc_fl.p, ix->F,
                                                     d_keys.p, d_vals.p, d_max.p);
          CU(cudaGetLastError());
        }
        uint32_t h_max[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        CU(cudaMemcpy(h_max, d_max.p, sizeof(h_max), cudaMemcpyDeviceToHost));
        for (uint32_t f = 0; f < ix->F; ++f) { ix->max_tf[f] = h_max[f]; ix->max_fl[f] = h_max[4 + f]; }
      }
      bool fits = true;
      for (uint32_t f = 0; f < ix->F; ++f) {
        ix->fl_bits[f] = bits_for((uint64_t)ix->max_fl[f] + 1);
        fits = fits && (((uint64_t)ix->max_tf[f] + 1) << ix->fl_bits[f]) <= 65536ull;
      }
      uint64_t min_table = 0;       // the BM25 table keeps whole rows of 1 << fl_bits entries in the narrow layout
      for (uint32_t f = 0; f < ix->F; ++f) min_table += 4ull << ix->fl_bits[f];
      fits = fits && min_table <= 8192;
      const char* e = std::getenv("PB_POSTING_LAYOUT");
      if (e && !std::strcmp(e, "narrow") && !fits) { pb::set_error("PB_POSTING_LAYOUT=narrow but a (tf, field length) pair does not fit 16 bits"); return PB_ERR_UNSUPPORTED; }
      ix->narrow = fits && !(e && !std::strcmp(e, "wide"));
      const uint32_t NC = 1 + 2 * ix->F;
      const uint64_t tiles = im->n_rows_padded / TILE_ROWS;
      if (log) {
        // part 2: stable radix sort by term ordinal (the log is in document order, so docs ascend inside a term), then
        // the tiles are written where they will be read
        ix->tile_words = ix->narrow ? TILE_ROWS + ix->F * (TILE_ROWS / 2) : NC * TILE_ROWS;
        CU(ix->post_blocks.ensure((size_t)(tiles + 1) * ix->tile_words + 1));      // zero-filled: pad rows, spare tile
        const uint64_t n = log->n_tuples;
        if (n) {
          const int end_bit = (int)bits_for(std::max<uint64_t>(im->n_terms, 2));
          size_t bytes = 0;
          DBuf<uint8_t> tmp;
          CU(cub::DeviceRadixSort::SortPairs(nullptr, bytes, d_keys.p, d_keys2.p, d_vals.p, d_vals2.p, (int64_t)n, 0, end_bit));
          CU(tmp.ensure(bytes));
          bytes = tmp.cap;
          CU(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, d_keys.p, d_keys2.p, d_vals.p, d_vals2.p, (int64_t)n, 0, end_bit));
          flat_gather_kernel<<<(unsigned)((n + 255) / 256), 256>>>(reinterpret_cast<const FlatTuple*>(d_log.p), d_vals2.p, n, d_doc_fl.p,
                                                                   ix->F, ix->narrow ? 1u : 0u, ix->tile_words, ix->fl_bits[0],
                                                                   ix->fl_bits[1], ix->fl_bits[2], ix->fl_bits[3], ix->post_blocks.p);
          CU(cudaGetLastError());
          CU(cudaDeviceSynchronize());
        }
      } else if (!ix->narrow) {
        ix->tile_words = NC * TILE_ROWS;
        CU(upload(ix->post_blocks, im->post_blocks, im->n_rows_padded * NC, (size_t)TILE_ROWS * NC));
      } else {
        ix->tile_words = TILE_ROWS + ix->F * (TILE_ROWS / 2);
        std::vector<uint32_t> nb((size_t)(tiles + 1) * ix->tile_words, 0u);      // + one zero tile behind the spare tile
        for (uint64_t t = 0; t < tiles; ++t) {
          const uint32_t* src = im->post_blocks + t * (uint64_t)NC * TILE_ROWS;
          uint32_t* dst = nb.data() + t * (uint64_t)ix->tile_words;
          std::memcpy(dst, src, TILE_ROWS * sizeof(uint32_t));
          uint16_t* codes = reinterpret_cast<uint16_t*>(dst + TILE_ROWS);
          for (uint32_t f = 0; f < ix->F; ++f)
            for (uint32_t r = 0; r < (uint32_t)TILE_ROWS; ++r)
              codes[f * TILE_ROWS + r] = (uint16_t)((src[(1 + f) * TILE_ROWS + r] << ix->fl_bits[f]) | src[(1 + ix->F + f) * TILE_ROWS + r]);
        }
        CU(upload(ix->post_blocks, nb.data(), nb.size(), 0));
      }
    }
    RC(index_build_compact(ix));
    RC(index_build_union(ix));
    // rank directories of the dense lists (>= n_docs / 32 rows, capped at 1 GB): built on the device
    {
      std::vector<uint32_t> tdir(im->n_terms + 1, 0u), dense;
      ix->dir_words = (uint32_t)((im->n_docs + 31) / 32 + 1);
      uint64_t min_rows = std::max<uint64_t>(4096, im->n_docs / 32);
      if (const char* e = std::getenv("PB_DIR_MIN_ROWS")) min_rows = std::max<long long>(1, atoll(e));   // tests: directories for every list
      const uint64_t max_dense = (1ull << 30) / ((uint64_t)ix->dir_words * sizeof(uint2));
      for (uint64_t t = 0; t < im->n_terms && dense.size() < max_dense; ++t)
        if (im->term_row_begin[t + 1] - im->term_row_begin[t] >= min_rows) { dense.push_back((uint32_t)t); tdir[t] = (uint32_t)dense.size(); }
      ix->n_dense = (uint32_t)dense.size();
      CU(upload(ix->term_dir, tdir.data(), tdir.size()));
      CU(ix->dir.ensure((size_t)ix->n_dense * ix->dir_words + 1));
      if (ix->n_dense) {
        DBuf<uint32_t> d_dense;
        CU(upload(d_dense, dense.data(), dense.size()));
        CU(cudaMemset(ix->dir.p, 0, (size_t)ix->n_dense * ix->dir_words * sizeof(uint2)));
        IndexView v = ix->view();
        dir_bits_kernel<<<dim3(64, std::min<uint32_t>(ix->n_dense, 1024)), 256>>>(v, d_dense.p, ix->n_dense, ix->dir.p);
        CU(cudaGetLastError());
        dir_rank_kernel<<<ix->n_dense, 1024>>>(ix->dir_words, ix->dir.p);
        CU(cudaGetLastError());
        CU(cudaDeviceSynchronize());
      }
    }
    ix->h_node_parent.assign(im->node_parent, im->node_parent + im->n_nodes);
    ix->h_node_char.assign(im->node_char, im->node_char + im->n_nodes);
    ix->h_term_node.assign(im->term_node, im->term_node + im->n_terms);
    ix->h_term_row_begin.assign(im->term_row_begin, im->term_row_begin + im->n_terms + 1);
    {
      std::vector<double> rc(1025, 0.0);
      for (int d = 1; d <= 1024; ++d) rc[d] = 1.0 / (double)d;     // correctly rounded by the host division
      CU(upload(ix->rcp, rc.data(), rc.size()));
    }
    // expansion boost by byte-length delta, bm25.rs:45-53 (libm log on the host)
    std::vector<double> eb(im->max_term_bytes + 2, 1.0);
    for (size_t d = 1; d < eb.size(); ++d) eb[d] = std::log(1.0 + (1.0 / (1.0 + (double)d)));   // (1 + explen) - qlen = 1 + d exactly
    CU(upload(ix->eb, eb.data(), eb.size()));
    RC(index_apply_live_state(ix, im->removed_bitmap, im->n_removed, im->n_live_docs, im->field_avg));
    *out = guard.release();
    return PB_OK;
  });
}

int pb_index_set_live_state(pb_index* ix, const uint32_t* removed_ords, uint64_t n_removed, uint64_t n_live_docs,
                            const double* field_avg) {
  if (!ix || !field_avg || (n_removed && !removed_ords)) { pb::set_error("pb_index_set_live_state: null argument"); return PB_ERR_INVALID; }
  PB_TRY({
    StateChange lk(ix);
    // with a delta segment the ordinals are those of the whole index: the main image takes the ones it covers
    const uint64_t n_all = ix->delta ? ix->delta->n_docs : ix->n_docs;
    std::vector<uint32_t> bm((ix->n_docs + 31) / 32 + 1, 0), bmd;
    if (ix->delta) bmd.assign((n_all + 31) / 32 + 1, 0);
    uint64_t distinct = 0, distinct_d = 0;
    for (uint64_t i = 0; i < n_removed; ++i) {
      uint32_t d = removed_ords[i];
      if (d >= n_all) { pb::set_error("pb_index_set_live_state: ordinal %u out of range", d); return PB_ERR_INVALID; }
      if (ix->delta) {
        if (!((bmd[d >> 5] >> (d & 31)) & 1u)) ++distinct_d;
        bmd[d >> 5] |= 1u << (d & 31);
      }
      if (d < ix->n_docs) {
        if (!((bm[d >> 5] >> (d & 31)) & 1u)) ++distinct;
        bm[d >> 5] |= 1u << (d & 31);
      }
    }
    RC(index_apply_live_state(ix, bm.data(), distinct, n_live_docs, field_avg));
    if (ix->delta) {
      RC(index_apply_live_state(ix->delta, bmd.data(), distinct_d, n_live_docs, field_avg));
      RC(segments_exchange_df(ix));
    }
    return PB_OK;
  });
}

int pb_index_set_df_extra(pb_index* ix, const uint64_t* df_extra, uint64_t n) {
  if (!ix || (n && !df_extra)) { pb::set_error("pb_index_set_df_extra: null argument"); return PB_ERR_INVALID; }
  if (n != 0 && n != ix->n_terms) { pb::set_error("pb_index_set_df_extra: %llu entries, the image has %llu terms", (ull)n, (ull)ix->n_terms); return PB_ERR_INVALID; }
  PB_TRY({
    StateChange lk(ix);
    CU(cudaSetDevice(ix->device));
    ix->h_df_extra.assign(df_extra, df_extra + n);
    ++ix->live_epoch;
    return index_upload_idf(ix);
  });
}

int pb_index_attach_delta(pb_index* ix, pb_index* delta, const uint32_t* main_term_ids, uint64_t n_main_terms,
                          const uint32_t* delta_term_ids, uint64_t n_delta_terms) {
  if (!ix) { pb::set_error("pb_index_attach_delta: null argument"); return PB_ERR_INVALID; }
  PB_TRY({
    StateChange lk(ix);
    if (delta) {
      if (delta == ix || delta->delta) { pb::set_error("pb_index_attach_delta: a delta segment cannot have one of its own"); return PB_ERR_INVALID; }
      if (delta->device != ix->device || delta->F != ix->F) { pb::set_error("pb_index_attach_delta: the segments must live on one device and have the same fields"); return PB_ERR_INVALID; }
      if (n_main_terms != ix->n_terms || n_delta_terms != delta->n_terms || (n_main_terms && !main_term_ids) || (n_delta_terms && !delta_term_ids)) {
        pb::set_error("pb_index_attach_delta: term id arrays must have one entry per term of each image"); return PB_ERR_INVALID;
      }
      if (delta->n_docs < ix->n_docs) { pb::set_error("pb_index_attach_delta: the delta image must cover the whole ordinal range"); return PB_ERR_INVALID; }
    }
    if (ix->delta && ix->delta != delta) pb_index_destroy(ix->delta);
    ix->delta = delta;
    if (delta) {
      ix->sid.assign(main_term_ids, main_term_ids + n_main_terms);
      delta->sid.assign(delta_term_ids, delta_term_ids + n_delta_terms);
    } else {
      ix->sid.clear();
    }
    return segments_exchange_df(ix);
  });
}

void pb_index_destroy(pb_index* ix) {
  if (!ix) return;
  cudaSetDevice(ix->device);
  if (ix->delta) pb_index_destroy(ix->delta);
  for (pb_batch* b : ix->pool_free) delete b;
  delete ix;
}

int pb_index_term_df_live(pb_index* ix, uint64_t* out, uint64_t cap) {
  if (!ix || !out) return PB_ERR_INVALID;
  if (cap < ix->n_terms) { pb::set_error("pb_index_term_df_live: need %llu entries", (ull)ix->n_terms); return PB_ERR_CAPACITY; }
  for (uint64_t t = 0; t < ix->n_terms; ++t) out[t] = ix->h_df_live[t];
  return PB_OK;
}

namespace {
__global__ void __launch_bounds__(256) read_bw_kernel(const uint4* __restrict__ p, uint64_t n, uint32_t passes, unsigned long long* sink) {
  const uint64_t i0 = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint32_t acc = 0;
  for (uint32_t pass = 0; pass < passes; ++pass) {     // several passes per launch: launch gaps do not count
    uint64_t i = i0;
    for (; i + 3 * stride < n; i += 4 * stride) {      // four independent 128-bit loads in flight per thread
      uint4 a = ldg_stream((const uint32_t*)(p + i)), b = ldg_stream((const uint32_t*)(p + i + stride));
      uint4 c = ldg_stream((const uint32_t*)(p + i + 2 * stride)), d = ldg_stream((const uint32_t*)(p + i + 3 * stride));
      acc += (a.x ^ b.y) + (c.z ^ d.w);
    }
    for (; i < n; i += stride) acc += ldg_stream((const uint32_t*)(p + i)).x;
  }
  if (acc == 0x9E3779B9u) atomicAdd(sink, 1ull);      // keeps the loads alive
}
}  // namespace

int pb_device_read_bandwidth(int device, uint64_t bytes, uint32_t iters, double* gb_per_s) {
  if (!gb_per_s || bytes < 4096 || iters == 0) { pb::set_error("pb_device_read_bandwidth: bad argument"); return PB_ERR_INVALID; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); pb::set_error("no CUDA device is available"); return PB_ERR_NO_DEVICE; }
  CU(cudaSetDevice(device));
  DBuf<uint4> buf;
  DBuf<ull> sink;
  const uint64_t n = bytes / sizeof(uint4);
  CU(buf.ensure(n));
  CU(sink.ensure(1));
  CU(cudaMemset(buf.p, 1, n * sizeof(uint4)));
  CU(cudaMemset(sink.p, 0, sizeof(ull)));
  struct Events {      // destroyed on every return path
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~Events() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
  } ev;
  CU(cudaEventCreate(&ev.e0)); CU(cudaEventCreate(&ev.e1));
  const int grid = g_sm_count(device) * 8;
  read_bw_kernel<<<grid, 256>>>(buf.p, n, 1, sink.p);
  CU(cudaEventRecord(ev.e0));
  read_bw_kernel<<<grid, 256>>>(buf.p, n, iters, sink.p);
  CU(cudaEventRecord(ev.e1));
  CU(cudaEventSynchronize(ev.e1));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, ev.e0, ev.e1));
  CU(cudaGetLastError());
  *gb_per_s = (double)n * sizeof(uint4) * iters / (ms * 1e-3) / 1e9;
  return PB_OK;
}

int pb_index_device_layout(pb_index* ix, pb_device_layout* out) {
  if (!ix || !out) { pb::set_error("pb_index_device_layout: null argument"); return PB_ERR_INVALID; }
  std::memset(out, 0, sizeof(*out));
  out->narrow = ix->narrow ? 1u : 0u;
  out->bytes_per_row = ix->narrow ? 4u + 2u * ix->F : 4u + 8u * ix->F;
  for (uint32_t f = 0; f < ix->F; ++f) out->fl_bits[f] = ix->fl_bits[f];
  out->posting_bytes = (ix->n_rows_padded / TILE_ROWS) * (uint64_t)ix->tile_words * 4ull;
  return PB_OK;
}

int pb_index_expand_term(pb_index* ix, const uint8_t* term, uint64_t term_len, uint8_t* out, uint64_t cap,
                         uint64_t* n_expansions, uint64_t* needed) {
  if (!ix || !n_expansions || !needed || (term_len && !term)) { pb::set_error("pb_index_expand_term: null argument"); return PB_ERR_INVALID; }
  if (ix->delta) { pb::set_error("pb_index_expand_term: the index has a delta segment (fold it in first)"); return PB_ERR_UNSUPPORTED; }
  PB_TRY({
    *n_expansions = 0; *needed = 0;
    if (!pb::utf8_valid(term, term_len)) { pb::set_error("pb_index_expand_term: invalid UTF-8"); return PB_ERR_INVALID; }
    ScratchLease lease(ix);
    RC(lease.acquire());
    pb_batch* b = lease.b;
    CU(cudaSetDevice(ix->device));
    uint64_t off[2] = {0, term_len};
    CU(b->term_bytes.ensure(term_len + 16)); CU(b->term_byte_off.ensure(3));
    CU(b->qt_lo.ensure(2)); CU(b->qt_hi.ensure(2)); CU(b->qt_len.ensure(2));
    if (term_len) CU(cudaMemcpyAsync(b->term_bytes.p, term, term_len, cudaMemcpyHostToDevice, b->stream));
    CU(cudaMemcpyAsync(b->term_byte_off.p, off, sizeof(off), cudaMemcpyHostToDevice, b->stream));
    b->loaded = false;
    descend_kernel<<<1, 32, 0, b->stream>>>(ix->view(), b->term_bytes.p, b->term_byte_off.p, 1, b->qt_lo.p, b->qt_hi.p, b->qt_len.p);
    CU(cudaGetLastError());
    uint32_t lo = 0, hi = 0;
    CU(cudaMemcpyAsync(&lo, b->qt_lo.p, 4, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaMemcpyAsync(&hi, b->qt_hi.p, 4, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    // the device kernel skips empty tokens (query.rs:35); expand_term("") itself expands the root
    if (term_len == 0) { lo = 0; hi = (uint32_t)ix->n_terms; }
    uint64_t pos = 0;
    std::string s;
    for (uint32_t t = lo; t < hi; ++t) {
      s.clear();
      std::vector<uint32_t> cps;
      for (uint32_t n = ix->h_term_node[t]; n != 0 && n != NONE; n = ix->h_node_parent[n]) cps.push_back(ix->h_node_char[n]);
      for (size_t i = cps.size(); i-- > 0;) {
        uint32_t c = cps[i];
        if (c < 0x80) s.push_back((char)c);
        else if (c < 0x800) { s.push_back((char)(0xC0 | (c >> 6))); s.push_back((char)(0x80 | (c & 0x3F))); }
        else if (c < 0x10000) { s.push_back((char)(0xE0 | (c >> 12))); s.push_back((char)(0x80 | ((c >> 6) & 0x3F))); s.push_back((char)(0x80 | (c & 0x3F))); }
        else { s.push_back((char)(0xF0 | (c >> 18))); s.push_back((char)(0x80 | ((c >> 12) & 0x3F))); s.push_back((char)(0x80 | ((c >> 6) & 0x3F))); s.push_back((char)(0x80 | (c & 0x3F))); }
      }
      if (t > lo) { if (pos < cap && out) out[pos] = '\n'; ++pos; }
      for (char c : s) { if (pos < cap && out) out[pos] = (uint8_t)c; ++pos; }
    }
    *n_expansions = hi - lo;
    *needed = pos;
    return PB_OK;
  });
}

int pb_batch_create(pb_index* ix, const pb_query_batch_desc* q, pb_batch** out) {
  if (!ix || !q || !out) { pb::set_error("pb_batch_create: null argument"); return PB_ERR_INVALID; }
  if (ix->delta) { pb::set_error("pb_batch_create: the index has a delta segment; a staged batch runs on one image (fold the delta in, or use pb_query_batch)"); return PB_ERR_UNSUPPORTED; }
  PB_TRY({
    pb_batch* b = nullptr;
    RC(batch_new(ix, &b));
    int rc = batch_load(b, q, 0);
    if (rc != PB_OK) { delete b; return rc; }
    *out = b;
    return PB_OK;
  });
}

int pb_batch_run(pb_batch* b) {
  if (!b) return PB_ERR_INVALID;
  PB_TRY({ return batch_run(b); });
}

int pb_batch_fetch(pb_batch* b, pb_query_results* out) {
  if (!b || !out) return PB_ERR_INVALID;
  PB_TRY({ return batch_fetch(b, out); });
}

void pb_batch_destroy(pb_batch* b) { delete b; }

int pb_batch_reload(pb_batch* b, const pb_query_batch_desc* q) {
  if (!b || !q) { pb::set_error("pb_batch_reload: null argument"); return PB_ERR_INVALID; }
  PB_TRY({ return batch_load(b, q, 0); });
}

int pb_batch_set_gather(pb_batch* b, pb_comm* comm, uint64_t slot_queries) {
  if (!b) { pb::set_error("pb_batch_set_gather: null argument"); return PB_ERR_INVALID; }
  PB_TRY({
    if (comm && pbg::comm_device(comm) != b->ix->device) { pb::set_error("pb_batch_set_gather: the communicator lives on device %d, the batch on device %d", pbg::comm_device(comm), b->ix->device); return PB_ERR_INVALID; }
    if (comm && slot_queries < b->Q) { pb::set_error("pb_batch_set_gather: slot of %llu queries is smaller than the batch (%llu)", (ull)slot_queries, (ull)b->Q); return PB_ERR_INVALID; }
    CU(cudaSetDevice(b->ix->device));
    CU(cudaStreamSynchronize(b->stream));
    b->comm = comm;
    b->want_slot = comm ? slot_queries : 0;
    b->ran = false; b->gathered = false;
    return batch_layout_results(b);
  });
}

int pb_batch_run_local(pb_batch* b) {
  if (!b) return PB_ERR_INVALID;
  PB_TRY({
    RC(batch_compute(b));
    return batch_finish(b);
  });
}

int pb_batch_gather(pb_batch* b) {
  if (!b) return PB_ERR_INVALID;
  if (!b->ran) { pb::set_error("pb_batch_gather: batch has not been run"); return PB_ERR_INVALID; }
  PB_TRY({ return batch_gather(b); });
}

int pb_batch_sync(pb_batch* b) {
  if (!b) return PB_ERR_INVALID;
  PB_TRY({
    CU(cudaSetDevice(b->ix->device));
    CU(cudaStreamSynchronize(b->stream));
    if (b->gather_pending) {
      float ms = 0;
      CU(cudaEventElapsedTime(&ms, b->ev[6], b->ev[7]));
      b->st.ms_gather = ms;
      b->gathered = true;
      b->gather_pending = false;
    }
    return PB_OK;
  });
}

int pb_batch_fetch_gathered(pb_batch* b, uint64_t n_total, pb_query_results* o) {
  if (!b || !o) return PB_ERR_INVALID;
  PB_TRY({
    RC(pb_batch_sync(b));
    if (!b->comm || !b->gathered) { pb::set_error("pb_batch_fetch_gathered: no gathered results (pb_batch_set_gather + pb_batch_run)"); return PB_ERR_INVALID; }
    const uint64_t world = (uint64_t)pbg::comm_world(b->comm), slot = b->slot;
    if (n_total > world * slot) { pb::set_error("pb_batch_fetch_gathered: %llu queries asked, the gather holds %llu", (ull)n_total, (ull)(world * slot)); return PB_ERR_INVALID; }
    const ResLayout L = res_layout(slot, b->k);
    cudaStream_t st = b->stream;
    const uint32_t k = b->k;
    for (uint64_t r = 0; r < world; ++r) {
      const uint64_t g0 = r * slot;
      if (g0 >= n_total) break;
      const uint64_t n = std::min<uint64_t>(slot, n_total - g0);
      const uint8_t* blk = b->gath.p + r * b->block_bytes;
      if (o->n_results) CU(cudaMemcpyAsync(o->n_results + g0, blk + L.n, n * 8, cudaMemcpyDeviceToHost, st));
      if (o->doc_digest) CU(cudaMemcpyAsync(o->doc_digest + g0, blk + L.dd, n * 8, cudaMemcpyDeviceToHost, st));
      if (o->score_digest) CU(cudaMemcpyAsync(o->score_digest + g0, blk + L.sd, n * 8, cudaMemcpyDeviceToHost, st));
      if (o->topk_n) CU(cudaMemcpyAsync(o->topk_n + g0, blk + L.tn, n * 4, cudaMemcpyDeviceToHost, st));
      if (k && o->topk_doc) CU(cudaMemcpyAsync(o->topk_doc + g0 * k, blk + L.td, n * k * 4, cudaMemcpyDeviceToHost, st));
      if (k && o->topk_score) CU(cudaMemcpyAsync(o->topk_score + g0 * k, blk + L.ts, n * k * 8, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    return PB_OK;
  });
}

int pb_batch_device_gathered(pb_batch* b, const void** block0, uint64_t* block_bytes, uint64_t* slot_queries) {
  if (!b || !block0 || !block_bytes || !slot_queries) return PB_ERR_INVALID;
  if (!b->comm || !b->gathered) { pb::set_error("pb_batch_device_gathered: no gathered results"); return PB_ERR_INVALID; }
  *block0 = b->gath.p; *block_bytes = b->block_bytes; *slot_queries = b->slot;
  return PB_OK;
}

int pb_batch_device_results(pb_batch* b, pb_query_results* o) {
  if (!b || !o) return PB_ERR_INVALID;
  if (!b->ran) { pb::set_error("pb_batch_device_results: batch has not been run"); return PB_ERR_INVALID; }
  o->n_results = (uint64_t*)b->n_results; o->doc_digest = (uint64_t*)b->doc_digest;
  o->score_digest = (uint64_t*)b->score_digest; o->topk_n = b->topk_n; o->topk_doc = b->topk_doc;
  o->topk_score = b->topk_score;
  return PB_OK;
}

int pb_batch_get_stats(const pb_batch* b, pb_batch_stats* out) {
  if (!b || !out) return PB_ERR_INVALID;
  *out = b->st;
  return PB_OK;
}

static int query_batch_one(pb_index* ix, const pb_query_batch_desc* q, pb_query_results* out) {
  PB_TRY({
    ScratchLease lease(ix);
    RC(lease.acquire());
    pb_batch* b = lease.b;
    // one call = upload + kernels + download with as few host round trips as the planning allows: the caller's
    // buffers stay valid for the whole call, so nothing waits for the uploads, and the result copies are
    // enqueued BEFORE the run's final synchronisation (small batches: one copy of the packed block into a pinned
    // staging buffer instead of six)
    RC(batch_load(b, q, 0, false));
    RC(batch_compute(b));
    const bool staged = b->Q && b->block_bytes <= (1u << 20);
    if (staged) {
      if (b->h_stage_bytes < b->block_bytes) {
        if (b->h_stage) cudaFreeHost(b->h_stage);
        b->h_stage = nullptr; b->h_stage_bytes = 0;
        CU(cudaMallocHost(&b->h_stage, b->block_bytes));
        b->h_stage_bytes = b->block_bytes;
      }
      CU(cudaMemcpyAsync(b->h_stage, b->res.p, b->block_bytes, cudaMemcpyDeviceToHost, b->stream));
    } else {
      RC(batch_fetch_enqueue(b, out));
    }
    RC(batch_finish(b));
    if (staged) {
      const ResLayout L = res_layout(b->slot, b->k);
      const uint8_t* h = static_cast<const uint8_t*>(b->h_stage);
      const uint64_t Q = b->Q, k = b->k;
      if (out->n_results) std::memcpy(out->n_results, h + L.n, Q * 8);
      if (out->doc_digest) std::memcpy(out->doc_digest, h + L.dd, Q * 8);
      if (out->score_digest) std::memcpy(out->score_digest, h + L.sd, Q * 8);
      if (out->topk_n) std::memcpy(out->topk_n, h + L.tn, Q * 4);
      if (k && out->topk_doc) std::memcpy(out->topk_doc, h + L.td, Q * k * 4);
      if (k && out->topk_score) std::memcpy(out->topk_score, h + L.ts, Q * k * 8);
    }
    return PB_OK;
  });
}

int pb_query_batch(pb_index* ix, const pb_query_batch_desc* q, pb_query_results* out) {
  if (!ix || !q || !out) { pb::set_error("pb_query_batch: null argument"); return PB_ERR_INVALID; }
  if (!ix->delta) return query_batch_one(ix, q, out);
  // Segmented index: both segments answer the batch; a document lives in exactly one of them, so counts and digests
  // add and the two (score desc, doc asc) top-k lists merge.
  PB_TRY({
    const uint64_t Q = q->n_queries, k = q->top_k, kk = std::max<uint64_t>(k, 1);
    std::vector<uint64_t> n[2], dd[2], sd[2];
    std::vector<uint32_t> tn[2], td[2];
    std::vector<double> ts[2];
    pb_index* seg[2] = {ix, ix->delta};
    for (int s2 = 0; s2 < 2; ++s2) {
      n[s2].assign(Q + 1, 0); dd[s2].assign(Q + 1, 0); sd[s2].assign(Q + 1, 0); tn[s2].assign(Q + 1, 0);
      td[s2].assign(Q * kk + 1, 0); ts[s2].assign(Q * kk + 1, 0.0);
      pb_query_results r = {n[s2].data(), dd[s2].data(), sd[s2].data(), tn[s2].data(), td[s2].data(), ts[s2].data()};
      RC(query_batch_one(seg[s2], q, &r));
    }
    for (uint64_t i = 0; i < Q; ++i) {
      if (out->n_results) out->n_results[i] = n[0][i] + n[1][i];
      if (out->doc_digest) out->doc_digest[i] = dd[0][i] + dd[1][i];
      if (out->score_digest) out->score_digest[i] = sd[0][i] + sd[1][i];
      uint32_t a = 0, b2 = 0, m = 0;
      const uint32_t na = tn[0][i], nb = tn[1][i];
      while (m < k && (a < na || b2 < nb)) {
        bool take_a = b2 >= nb;
        if (a < na && b2 < nb) {
          const double sa = ts[0][i * kk + a], sb = ts[1][i * kk + b2];
          take_a = sa > sb || (sa == sb && td[0][i * kk + a] < td[1][i * kk + b2]);
        }
        const int s2 = take_a ? 0 : 1;
        const uint64_t src = i * kk + (take_a ? a++ : b2++);
        if (out->topk_doc) out->topk_doc[i * k + m] = td[s2][src];
        if (out->topk_score) out->topk_score[i * k + m] = ts[s2][src];
        ++m;
      }
      if (out->topk_n) out->topk_n[i] = m;
    }
    return PB_OK;
  });
}

int pb_index_last_stats(pb_index* ix, pb_batch_stats* out) {
  if (!ix || !out) return PB_ERR_INVALID;
  std::lock_guard<std::mutex> lk(ix->pool_mu);
  if (!ix->has_last) return PB_ERR_INVALID;
  *out = ix->last_st;
  return PB_OK;
}

static int query_full_one(pb_index* ix, const pb_query_batch_desc* q, uint64_t cap, uint32_t* out_query, uint32_t* out_doc,
                          double* out_score, uint64_t* n_total) {
  PB_TRY({
    ScratchLease lease(ix);
    RC(lease.acquire());
    pb_batch* b = lease.b;
    pb_query_batch_desc d = *q;
    RC(batch_load(b, &d, std::max<uint64_t>(cap, 1)));
    int rc = batch_run(b);
    b->full_cap = 0;
    RC(rc);
    ull total = 0;
    CU(cudaMemcpy(&total, b->full_count.p, sizeof(ull), cudaMemcpyDeviceToHost));
    *n_total = total;
    if (total > cap) { pb::set_error("pb_query_full: %llu results, capacity %llu", total, (ull)cap); return PB_ERR_CAPACITY; }
    if (total) {
      if (!out_query || !out_doc || !out_score) { pb::set_error("pb_query_full: null output"); return PB_ERR_INVALID; }
      CU(cudaMemcpy(out_query, b->full_q.p, total * sizeof(uint32_t), cudaMemcpyDeviceToHost));
      CU(cudaMemcpy(out_doc, b->full_doc.p, total * sizeof(uint32_t), cudaMemcpyDeviceToHost));
      CU(cudaMemcpy(out_score, b->full_score.p, total * sizeof(double), cudaMemcpyDeviceToHost));
    }
    return PB_OK;
  });
}

int pb_query_full(pb_index* ix, const pb_query_batch_desc* q, uint64_t cap, uint32_t* out_query, uint32_t* out_doc,
                  double* out_score, uint64_t* n_total) {
  if (!ix || !q || !n_total) { pb::set_error("pb_query_full: null argument"); return PB_ERR_INVALID; }
  if (!ix->delta) return query_full_one(ix, q, cap, out_query, out_doc, out_score, n_total);
  // segmented index: the result sets of the two segments are disjoint by document: concatenated
  uint64_t n0 = 0, n1 = 0;
  int rc0 = query_full_one(ix, q, cap, out_query, out_doc, out_score, &n0);
  if (rc0 != PB_OK && rc0 != PB_ERR_CAPACITY) return rc0;
  const uint64_t used = rc0 == PB_OK ? n0 : cap;
  int rc1 = query_full_one(ix->delta, q, cap - used, out_query ? out_query + used : nullptr, out_doc ? out_doc + used : nullptr,
                           out_score ? out_score + used : nullptr, &n1);
  if (rc1 != PB_OK && rc1 != PB_ERR_CAPACITY) return rc1;
  *n_total = n0 + n1;
  if (rc0 == PB_ERR_CAPACITY || rc1 == PB_ERR_CAPACITY) {
    pb::set_error("pb_query_full: %llu results, capacity %llu", (ull)(n0 + n1), (ull)cap);
    return PB_ERR_CAPACITY;
  }
  return PB_OK;
}

void* pb_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void pb_host_free(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"
