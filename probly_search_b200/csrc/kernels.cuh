// Device code of the query hot path (sm_100a).  Everything the reference does per query in
// src/query.rs:21-164 and src/score/default/{bm25,zero_to_one}.rs happens in these kernels.
//
//   descend_kernel     find_inverted_index_node (index.rs:300-337) over the CSR trie; because
//                      nodes/terms are numbered in DFS pre-order the whole of expand_term
//                      (query.rs:109-147) collapses to the term range [lo, hi) of the node.
//   plan_* / g*_kernel per query: how many live expanded lists, which class (single list ->
//                      streamed directly; several lists -> side path: primary / exact scheme),
//                      segment descriptors, doc-range bins, row-mask / bitmap words.
//   live_df_kernel     count_documents (index.rs:282-297) for every term at once.
//   dir_*_kernel       rank directories of the dense posting lists (built once per index).
//   score_kernel       THE hot loop (query.rs:61-89): posting rows -> removed mask -> BM25 /
//                      zero-to-one -> fused count / digest / top-k, or diversion to the side path.
//   mark_kernel        side path, pass 1: which rows of a multi-list query belong to docs hit by
//                      several (query term, expansion) events (row masks / doc bitmaps), and how
//                      many records every doc-range bin will receive.
//   binfold_kernel     side path, pass 3: max_score_merger (query.rs:150-164) and
//   (fold_kernel)      ZeroToOne::finalize (zero_to_one.rs:84-126) on the diverted events, per doc,
//                      in (query term, expansion) order; fold_kernel is the sorted fallback.
//   finalize_kernel    merges per-warp partial top-k lists (query.rs:97-105: result + sort).
//
// f64 arithmetic is done with __d*_rn intrinsics in the reference's operation order, so no FMA
// contraction can change a bit (rustc never contracts).  Both logarithms of BM25 are per term
// and are tabulated on the host with libm (engine.cu), never computed on the device.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

namespace pbk {

constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr int TILE_ROWS = 128;          // 32 lanes x 4 rows: one 512 B line group per column
constexpr int WARPS_PER_CTA = 8;
constexpr int CTA_THREADS = WARPS_PER_CTA * 32;
#ifndef PB_WINDOW_MAX
#define PB_WINDOW_MAX 32u      // records a bin-fold window holds; larger bins take the sorted fallback (tests lower it)
#endif

// How a posting list of a multi-list query is treated by the scoring kernel:
//   PRIMARY / SECONDARY ("primary scheme", few secondary rows): every secondary row is diverted to
//     the fold; while marking, each secondary doc is looked up in the primary (largest) list by binary
//     search and, when present, sets that ROW's bit in the query's row mask (1 bit per primary row,
//     tile-aligned), so the scoring pass learns which primary rows to divert from one coalesced
//     16-byte load per tile at an address that does not depend on the posting data.
//   MULTI ("exact scheme", many secondary rows): a marking pass over ALL lists finds the docs hit
//     by >= 2 (query term, expansion) events in per-query doc bitmaps; every list diverts exactly
//     those rows.
enum SegMode : uint8_t { MODE_DIRECT = 0, MODE_PRIMARY = 1, MODE_SECONDARY = 2, MODE_MULTI = 3 };

// One posting list walked for one (query term, expanded term): the unit query.rs:38-91 iterates.
struct __align__(16) Seg {
  uint64_t row_begin;
  uint32_t n_rows;
  uint32_t q;          // query index in the batch
  uint32_t term;       // expanded term ordinal (DFS order)
  uint32_t qlen;       // UTF-8 byte length of the query term
  uint16_t qti;        // query_term_index (query.rs:34)
  uint8_t mode;
  uint8_t pad;
  uint32_t slot;       // rank of the query among the class-G queries of its side-path round (sorted-fallback key)
};
static_assert(sizeof(Seg) == 32, "Seg layout");

struct IndexView {
  const uint32_t* node_edge_begin;
  const uint32_t* node_term_lo;
  const uint32_t* node_term_hi;
  const uint32_t* edge_char;
  const uint32_t* edge_child;
  const uint64_t* term_row_begin;
  const uint32_t* term_byte_len;
  // Posting columns in HBM, tile-blocked (one contiguous block per 128 rows):
  //   wide   (any tf / field length):  [tile][doc u32 x128][tf0 u32 x128]..[fl(F-1) u32 x128]   = 4 + 8F bytes / row
  //   narrow (chosen at pb_index_create when, per field, (max tf + 1) << fl_bits <= 65536 with
  //          fl_bits = bits of the largest field length): one u16 code = tf << fl_bits | fl per field
  //          [tile][doc u32 x128][code0 u16 x128]..[code(F-1) u16 x128]                          = 4 + 2F bytes / row
  //          The code IS the index into the field's BM25 table (row stride 1 << fl_bits).
  const uint32_t* post_blocks;
  uint32_t tile_words;            // u32 words per tile: 128 (1 + 2F) wide, 128 + 64F narrow
  uint32_t narrow;
  // Compact copy of the narrow tiles for the single-list stream (delta-coded doc ordinals): the doc
  // column of a tile as u16 offsets from the tile's smallest doc, where that span fits 16 bits
  //          [tile][doc - base u16 x128][code0 u16 x128]..[code(F-1) u16 x128]                   = 2 + 2F bytes / row
  // same tile index as post_blocks (stride 64 (1 + F) words), `cbase[tile]` = the base or NONE (tile not compact);
  // `term_compact[t]` = every interior tile of term t's list is compact, i.e. the list may be streamed from here.
  // Built on the device at pb_index_create on request (PB_POSTING_COMPACT=1; measured 5-6 % slower than the u32 doc
  // column although it reads 25 % fewer bytes: the stream is issue-bound); null otherwise.
  // Only the streaming loop of class S reads it: marking, directories, the union image and random access keep
  // reading post_blocks.
  const uint32_t* cpost;
  const uint32_t* cbase;
  const uint8_t* term_compact;
  uint32_t fl_bits[4];
  const uint32_t* removed;        // bitmap, bit set = doc not live
  // The same fact per posting ROW (bit r % 128 of the 4 words of tile r / 128; rebuilt by pb_index_set_live_state when
  // docs are removed, null while none is): the scoring loop takes this lane's word with the tile - one load whose
  // address does not depend on posting data - instead of four dependent probes of `removed` per lane and tile.
  const uint32_t* row_dead;
  const uint64_t* term_df_live;
  const uint32_t* term_live_rows; // rows of the term whose doc is live
  const uint32_t* live_prefix;    // [n_terms+1] number of terms with df_live > 0 before t
  const uint64_t* liverows_prefix;// [n_terms+1] rows of live terms before t
  const double* term_idf;         // bm25.rs:56, host libm
  const double* eb;               // bm25.rs:45-53 by byte-length delta, host libm
  // Rank directory of the DENSE posting lists (rows >= n_docs / 32), built on the device at
  // pb_index_create: per 32-doc word {rows of the list before this word, bits of the docs present}.
  // row of doc d in the list = row_begin + dir.x + popc(dir.y & ((1 << (d & 31)) - 1)): the marking
  // pass finds a secondary doc in a dense primary list with ONE load instead of a binary search.
  const uint2* dir;               // [n_dense][dir_words]
  const uint32_t* term_dir;       // [n_terms] 1 + index of the term's directory, 0 = none
  uint32_t dir_words;
  const double* rcp;              // [1025] RN(1 / d), d = 1..1024 (PB_Z2O_RCP experiment; host-computed)
  uint32_t rcp_ok;                // every term is at most 255 bytes long: the proven domain of that experiment
  uint32_t n_terms;
  uint32_t n_docs;
  uint32_t num_fields;
  uint32_t has_removed;
};

// Random access into the tile-blocked columns (side path, live_df, marking): column 0 = doc,
// 1 + f = tf[f], 1 + F + f = fl[f].  Layout-aware at run time (these paths are not the hot loop).
__device__ __forceinline__ const uint32_t* tile_ptr(const IndexView& ix, uint64_t tile) {
  return ix.post_blocks + tile * (uint64_t)ix.tile_words;
}
__device__ __forceinline__ uint32_t row_doc(const IndexView& ix, uint64_t r) {
  return tile_ptr(ix, r / TILE_ROWS)[r % TILE_ROWS];
}
__device__ __forceinline__ uint32_t row_col(const IndexView& ix, uint64_t r, int c) {   // c = 0 .. 2F-1 (tf.., fl..)
  const uint32_t* t = tile_ptr(ix, r / TILE_ROWS) + TILE_ROWS;
  if (ix.narrow) {
    const uint32_t nf = ix.num_fields, f = (uint32_t)c < nf ? (uint32_t)c : (uint32_t)c - nf;
    const uint32_t code = reinterpret_cast<const uint16_t*>(t)[f * TILE_ROWS + (r % TILE_ROWS)];
    return (uint32_t)c < nf ? code >> ix.fl_bits[f] : code & ((1u << ix.fl_bits[f]) - 1u);
  }
  return t[c * TILE_ROWS + (r % TILE_ROWS)];
}
template <int F> __device__ __forceinline__ uint32_t row_tf(const IndexView& ix, uint64_t r, int f) { return row_col(ix, r, f); }
template <int F> __device__ __forceinline__ uint32_t row_fl(const IndexView& ix, uint64_t r, int f) { return row_col(ix, r, F + f); }

struct Outputs {
  unsigned long long* n_results;
  unsigned long long* doc_digest;
  unsigned long long* score_digest;
  uint32_t* topk_n;
  uint32_t* topk_doc;
  double* topk_score;
  uint32_t k;
  // partial top-k lists (queries whose rows were handled by more than one warp / kernel)
  uint32_t* part_head;   // [n_queries]
  uint32_t* part_next;
  uint32_t* part_n;
  uint32_t* part_doc;    // [part_cap * k]
  double* part_score;
  uint32_t* part_count;
  uint32_t part_cap;
  // full result capture (pb_query_full)
  uint32_t* full_q;
  uint32_t* full_doc;
  double* full_score;
  unsigned long long* full_count;
  unsigned long long full_cap;
  uint32_t* error_flag;  // bit 0: partial list overflow, bit 1: record buffer overflow
  unsigned long long* results_total;   // sum of n_results over the batch
};

enum StatSlot { ST_ROWS_STREAMED = 0, ST_ROWS_SCORED, ST_POINTER_VISITS, ST_ROWS_DIVERTED, ST_ROWS_COMPACT, ST_COUNT };

struct ScoreParams {
  IndexView ix;
  Outputs out;
  const Seg* segs;
  const uint64_t* tile_off;      // exclusive prefix of tiles per segment, [n_segs + 1], absolute
  uint32_t seg_begin, seg_end;   // segment range of this launch
  uint64_t tile_begin, tile_end; // = tile_off[seg_begin], tile_off[seg_end]
  const uint64_t* query_term_off;// query_terms_len = off[q+1]-off[q] (query.rs:32)
  // BM25
  double k1, b, one_minus_b, k1_plus_1;
  double boost[4];               // fields_boost, with 1.0 for the fields whose boost is folded into the table (tab_scale)
  // A boost that is +-2^k multiplies exactly, and exact scaling commutes with every rounding of the score chain
  // ((tf' * idf) * boost) * eb  ==  ((tf' * boost) * idf) * eb  bit for bit - so the host folds it into the table of
  // saturated tf (and bm25_tf_slow applies it to values outside the table) and the loop skips that multiplication.
  double tab_scale[4];
  double avg[4];
  const double* tab;             // [F][tfcap][flcap] saturated tf (bm25.rs:78-82), host-computed
  uint32_t tab_tfcap[4], tab_flcap[4], tab_off[4], tab_total;
  uint32_t tab_full;             // the table covers every (tf, fl) present in the index
  // Shared-memory copy of the table: every entry is replicated 2^tab_rep_shift times, copy c of entry i
  // at byte (i << tab_rep_shift | c) * 8, and lane l reads copy l & (rep - 1).  With 16 copies the
  // 16 lanes of an LDS.64 phase always hit 16 different bank pairs: no bank conflicts whatever the
  // (tf, fl) values are.  tab_stride = 8 << tab_rep_shift bytes, tab_boff[f] = tab_off[f] * tab_stride.
  uint32_t tab_rep_shift, tab_stride, tab_boff[4];
  uint32_t boosts_all_one;       // every fields_boost is exactly 1.0
  uint32_t l2_prefetch;          // the posting image is far larger than L2 (streams from HBM): prefetch tiles ahead into L2
  // side path
  uint32_t* bitmap;              // pool of the round: query q's words start at q_bmoff[q] - round_bm0
  const unsigned long long* q_bmoff;   // [n_queries + 1] exclusive prefix of the per-query word counts
  unsigned long long round_bm0;
  const unsigned long long* q_prim;    // primary scheme: 0xFFFFFFFF - (low word) = segment index of the primary list
  unsigned long long* xtiles;          // tiles of exact-scheme lists walked by the marking pass (clear heuristic)
  uint32_t bitmap_words;         // words of an exact-scheme query (summary + two doc bit planes)
  uint32_t bitmap_sum_words;     // leading summary words of a slot
  uint32_t bitmap_doc_words;     // words of one per-doc bit plane (a slot has two)
  unsigned long long* xcount;    // exact-scheme rows the scoring pass will divert (counted while marking)
  // Side-path records are written straight into per-query doc-range BINS (no global sort): query q
  // owns bins [q_binoff[q], q_binoff[q+1]) and doc d falls into bin q_binoff[q] + (d >> q_shift[q]).
  // The marking pass counts each bin's capacity, a prefix sum gives bin_off, the scoring pass fills.
  const unsigned long long* q_binoff;
  const uint8_t* q_shift;
  const unsigned long long* q_gsegoff;  // first segment of each query (event order = seg - q_gsegoff[q])
  unsigned long long round_bin0;        // q_binoff of the round's first query
  uint32_t n_bins;                      // bins of the round
  uint32_t* bin_count;                  // [n_bins + 1] capacity (marking pass)
  uint32_t* bin_off;                    // [n_bins + 1] exclusive prefix of bin_count
  uint32_t* bin_cursor;                 // [n_bins] records written so far
  uint4* rec;                           // fat records {doc, segment, payload lo, payload hi}
  // legacy (sorted) records: only for bins that overflow a warp window
  unsigned long long* rec_key;
  unsigned long long* rec_val;
  uint32_t* rec_count;
  uint32_t rec_cap;
  uint32_t doc_bits;
  unsigned long long* stats;     // [ST_COUNT]
};

// ------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------
// Posting tiles are streamed once per segment; how they travel is a tuning knob:
//   PB_LDPOLICY 0: ld.global.nc.L1::no_allocate (pure stream)   1: ld.global.nc (allocates in L1)
#ifndef PB_LDPOLICY
#define PB_LDPOLICY 0
#endif
#ifndef PB_L2_AHEAD
#define PB_L2_AHEAD 6
#endif
__device__ __forceinline__ uint4 ldg_stream(const uint32_t* p) {
  uint4 r;
#if PB_LDPOLICY == 0
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
#else
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
#endif
  return r;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Digest terms (include/probly_b200.h "Digests"): cheap on purpose, they run once per result.
//   a(d) = (u32)(d + 1) * 0x9E3779B1            doc_digest   += (u64)a * a
//   y    = lo(s) ^ hi(s) * 0x85EBCA77 ^ a(d)    score_digest += (u64)y * y
// Each sum is one IMAD.WIDE.U32 into the 64-bit accumulator.
__device__ __forceinline__ uint32_t doc_mix(uint32_t doc) { return (doc + 1u) * 0x9E3779B1u; }
__device__ __forceinline__ uint32_t score_mix(uint32_t a, double s) {
  uint32_t lo = (uint32_t)__double2loint(s), hi = (uint32_t)__double2hiint(s);
  return lo ^ (hi * 0x85EBCA77u) ^ a;
}
__device__ __forceinline__ uint64_t sq64(uint32_t v) { return (uint64_t)v * v; }

__device__ __forceinline__ bool better(double as, uint32_t ad, double bs, uint32_t bd) {
  return as > bs || (as == bs && ad < bd);     // (score desc, doc asc), src/lib.rs:54-58
}

__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
  uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src);
  uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
  return (uint64_t(hi) << 32) | lo;
}
__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, o);
    uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), o);
    v += (uint64_t(hi) << 32) | lo;
  }
  return v;
}

// BM25 saturated term frequency, bm25.rs:78-82, in the reference's operation order.
__device__ __forceinline__ double bm25_tf_slow(const ScoreParams& P, uint32_t tf, uint32_t fl, int f) {
  double tfd = (double)tf;
  double num = __dmul_rn(P.k1_plus_1, tfd);
  double ratio = __ddiv_rn((double)fl, P.avg[f]);
  double inner = __dadd_rn(P.one_minus_b, __dmul_rn(P.b, ratio));
  double den = __dadd_rn(__dmul_rn(P.k1, inner), tfd);
  return __dmul_rn(__ddiv_rn(num, den), P.tab_scale[f]);     // 1.0 unless a power-of-two boost is folded in (exact)
}

// zero_to_one.rs:72 — 1 - |explen - qlen| / explen  (byte lengths)
__device__ __forceinline__ double z2o_term_score(uint32_t explen, uint32_t qlen) {
  double e = (double)explen, q = (double)qlen;
  return __dsub_rn(1.0, __ddiv_rn(fabs(__dsub_rn(e, q)), e));
}
// PB_Z2O_RCP = 1 (EXPERIMENT, default off, not yet run on a GPU): both divisions of the ZeroToOne entry
// below become  y = RN(1/d) from a 1025-entry table,  q = RN(x y),  r = fma(-q, d, x),  q' = fma(r, y, q).
// scripts/prove_z2o_rcp.c compares q' with the real quotient, bit for bit, for EVERY (term score, tf, m)
// the guard lets through (term byte lengths <= 255, tf <= 64, m <= 1024): 2.1e9 cases, 0 mismatches.
#ifndef PB_Z2O_RCP
#define PB_Z2O_RCP 0
#endif
__device__ __forceinline__ double div_small_int_rcp(double x, uint32_t d, const double* __restrict__ rcp) {
  const double y = __ldg(&rcp[d]);
  const double q = __dmul_rn(x, y);
  const double r = __fma_rn(-q, (double)d, x);
  return __fma_rn(r, y, q);
}
// zero_to_one.rs:117-120 — min(s/tf, 1) * tf / max(field_length, query_terms_len)
__device__ __forceinline__ double z2o_entry(const IndexView& ix, double s, uint32_t tf, uint32_t fl, uint32_t qtl) {
  const uint32_t m = max(fl, qtl);
#if PB_Z2O_RCP
  if (ix.rcp_ok && tf <= 64u && m <= 1024u) {
    const double v = __dmul_rn(fmin(div_small_int_rcp(s, tf, ix.rcp), 1.0), (double)tf);
    return div_small_int_rcp(v, m, ix.rcp);
  }
#endif
  double tfd = (double)tf;
  double v = __dmul_rn(fmin(__ddiv_rn(s, tfd), 1.0), tfd);
  return __ddiv_rn(v, (double)m);
}

// ------------------------------------------------------------------------------------------
// Per-warp result accumulator: count, digests, and a top-32 kept sorted across the lanes
// (lane i holds the i-th best).  One accumulator follows one query at a time.
// ------------------------------------------------------------------------------------------
struct WarpAcc {
  uint32_t q;
  uint32_t cnt;          // per lane
  uint64_t dd, sd;       // per lane
  double ts; uint32_t td;
  double thr_s; uint32_t thr_d;

  __device__ __forceinline__ void reset(uint32_t nq) {
    q = nq; cnt = 0; dd = 0; sd = 0;
    ts = -1.0; td = NONE; thr_s = -1.0; thr_d = NONE;
  }

  __device__ __forceinline__ void insert_candidates(bool c, uint32_t doc, double s, int lane, int k) {
    uint32_t m = __ballot_sync(0xffffffffu, c);
    while (m) {
      int l = __ffs(m) - 1;
      m &= m - 1;
      double cs = __shfl_sync(0xffffffffu, s, l);
      uint32_t cd = __shfl_sync(0xffffffffu, doc, l);
      if (!better(cs, cd, thr_s, thr_d)) continue;     // warp-uniform
      int pos = __popc(__ballot_sync(0xffffffffu, better(ts, td, cs, cd)));
      double us = __shfl_up_sync(0xffffffffu, ts, 1);
      uint32_t ud = __shfl_up_sync(0xffffffffu, td, 1);
      if (lane > pos) { ts = us; td = ud; }
      else if (lane == pos) { ts = cs; td = cd; }
      thr_s = __shfl_sync(0xffffffffu, ts, k - 1);
      thr_d = __shfl_sync(0xffffffffu, td, k - 1);
    }
  }

  __device__ __forceinline__ void add(const Outputs& o, bool valid, uint32_t doc, double s, int lane) {
    if (valid) {
      ++cnt;
      uint32_t a = doc_mix(doc);
      dd += sq64(a);
      sd += sq64(score_mix(a, s));
    }
    if (o.full_q) capture(o, valid, doc, s, lane);
    if (o.k) insert_candidates(valid && better(s, doc, thr_s, thr_d), doc, s, lane, (int)o.k);
  }

  // add() for scores that are >= +0.0 (ZeroToOne): the top-k pre-test compares bit patterns as signed integers
  // (doubles of one sign order like their bits; the threshold is -1.0 or such a score) instead of f64 compares.
  __device__ __forceinline__ void add_nonneg(const Outputs& o, bool valid, uint32_t doc, double s, int lane) {
    if (valid) {
      ++cnt;
      uint32_t a = doc_mix(doc);
      dd += sq64(a);
      sd += sq64(score_mix(a, s));
    }
    if (o.full_q) capture(o, valid, doc, s, lane);
    if (o.k) {
      const long long sb = __double_as_longlong(s), tb = __double_as_longlong(thr_s);
      const bool cand = valid && (sb > tb || (sb == tb && doc < thr_d));
      if (__any_sync(0xffffffffu, cand)) insert_candidates(cand, doc, s, lane, (int)o.k);
    }
  }

  // Four rows per lane at once (the scoring kernel's tile shape).  `some` = 4-bit mask of rows
  // that produced a result.  The top-k structure is only touched when some lane holds a score
  // that reaches the current k-th best.
  // TIES: scores that tie with the k-th best are common (ZeroToOne): the pre-test is the exact comparison, so a
  // tie with a larger doc ordinal does not enter the insertion loop.
  template <bool CAPTURE, bool TIES = false>
  __device__ __forceinline__ void add4(const Outputs& o, uint32_t some, const uint32_t (&doc)[4],
                                       const double (&sc)[4], int lane) {
    bool hit = false;
    if (TIES) {
      cnt += __popc(some);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t m = 0u - ((some >> j) & 1u);
        const uint32_t a = doc_mix(doc[j]);
        dd += sq64(a & m);
        sd += sq64(score_mix(a, sc[j]) & m);
        // scores here are >= +0.0 and the threshold is -1.0 or such a score: doubles of one sign order like their
        // bit patterns read as signed integers, so the exact (score desc, doc asc) test needs no f64 compare
        const long long sb = __double_as_longlong(sc[j]), tb = __double_as_longlong(thr_s);
        hit |= (m != 0u) && (sb > tb || (sb == tb && doc[j] < thr_d));
      }
    } else if (__all_sync(0xffffffffu, some == 0xFu)) {            // the common case: every row produced a result
      cnt += 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t a = doc_mix(doc[j]);
        dd += sq64(a);
        sd += sq64(score_mix(a, sc[j]));
      }
      // Some(score) is > 0 (bm25.rs:89) or >= +0 (zero_to_one): positive doubles order like their high
      // words, so "some score may reach the k-th best" is one integer compare (a superset; exact test below)
      const int hmax = max(max(__double2hiint(sc[0]), __double2hiint(sc[1])), max(__double2hiint(sc[2]), __double2hiint(sc[3])));
      hit = hmax >= __double2hiint(thr_s);
    } else {
      cnt += __popc(some);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t m = 0u - ((some >> j) & 1u);          // all ones when the row produced a result
        const uint32_t a = doc_mix(doc[j]);
        dd += sq64(a & m);
        sd += sq64(score_mix(a, sc[j]) & m);
        hit |= (m != 0u) && (sc[j] >= thr_s);
      }
    }
    if (CAPTURE && o.full_q) {
#pragma unroll
      for (int j = 0; j < 4; ++j) capture(o, (some >> j) & 1u, doc[j], sc[j], lane);
    }
    if (o.k && __any_sync(0xffffffffu, hit)) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        insert_candidates(((some >> j) & 1u) && better(sc[j], doc[j], thr_s, thr_d), doc[j], sc[j], lane, (int)o.k);
    }
  }

  __device__ __forceinline__ void capture(const Outputs& o, bool valid, uint32_t doc, double s, int lane) {
    uint32_t m = __ballot_sync(0xffffffffu, valid);
    if (m) {
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(o.full_count, (unsigned long long)__popc(m));
      base = shfl_u64(base, 0);
      unsigned long long pos = base + __popc(m & ((1u << lane) - 1u));
      if (valid && pos < o.full_cap) { o.full_q[pos] = q; o.full_doc[pos] = doc; o.full_score[pos] = s; }
    }
  }

  // owned: this warp saw every row of the query, so it may write the final top-k itself.
  __device__ __forceinline__ void flush(const Outputs& o, bool owned, int lane) {
    uint32_t total = cnt;
#pragma unroll
    for (int of = 16; of > 0; of >>= 1) total += __shfl_xor_sync(0xffffffffu, total, of);
    if (total == 0) return;
    uint64_t tdd = warp_sum_u64(dd), tsd = warp_sum_u64(sd);
    if (lane == 0) {
      atomicAdd(&o.n_results[q], (unsigned long long)total);
      atomicAdd(&o.doc_digest[q], (unsigned long long)tdd);
      atomicAdd(&o.score_digest[q], (unsigned long long)tsd);
      atomicAdd(o.results_total, (unsigned long long)total);
    }
    if (o.k == 0) return;
    uint32_t ntop = min(min(total, 32u), o.k);
    if (owned) {
      if (lane < (int)ntop) {
        o.topk_doc[(size_t)q * o.k + lane] = td;
        o.topk_score[(size_t)q * o.k + lane] = ts;
      }
      if (lane == 0) o.topk_n[q] = ntop;
    } else {
      uint32_t slot = 0;
      if (lane == 0) slot = atomicAdd(o.part_count, 1u);
      slot = __shfl_sync(0xffffffffu, slot, 0);
      if (slot >= o.part_cap) {
        if (lane == 0) atomicOr(o.error_flag, 1u);
        return;
      }
      if (lane < (int)ntop) {
        o.part_doc[(size_t)slot * o.k + lane] = td;
        o.part_score[(size_t)slot * o.k + lane] = ts;
      }
      if (lane == 0) {
        o.part_n[slot] = ntop;
        o.part_next[slot] = atomicExch(&o.part_head[q], slot);
      }
    }
  }
};

// count_documents (index.rs:282-297) for every term: live occurrence count
// df_live(t) = sum over the term's rows whose doc is live of sum_x tf[x]  (SURVEY §3.4 rule 2).
// Work is cut by ROWS, not by terms (posting lists span 1 ... 8.6e5 rows: one warp per term left the longest list
// on a single warp - 23 ms at 1 M docs): a warp owns a chunk of 2048 consecutive rows, follows the term boundaries as
// it walks (rows are term-major) and adds its partial sums with atomics; df_live / live_rows must be zero on entry.
constexpr uint32_t LIVE_DF_CHUNK = 2048;
template <int F>
__global__ void live_df_kernel(IndexView ix, unsigned long long* __restrict__ df_live,
                               uint32_t* __restrict__ live_rows) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  if (ix.n_terms == 0) return;
  const uint64_t n_rows = ix.term_row_begin[ix.n_terms];
  for (uint64_t c0 = warp * LIVE_DF_CHUNK; c0 < n_rows; c0 += nwarps * LIVE_DF_CHUNK) {
    const uint64_t c1 = min(c0 + (uint64_t)LIVE_DF_CHUNK, n_rows);
    // term of the chunk's first row: largest t with term_row_begin[t] <= c0
    uint32_t t = 0;
    {
      uint32_t lo = 0, hi = ix.n_terms;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (ix.term_row_begin[mid + 1] <= c0) lo = mid + 1; else hi = mid;
      }
      t = lo;
    }
    unsigned long long s = 0;      // this lane's partial sums for term t
    uint32_t n = 0;
    for (uint64_t r0 = c0; r0 < c1; r0 += 32) {
      const uint64_t r = r0 + lane;
      const uint64_t rend = min(r0 + 32, c1);
      uint32_t tf_sum = 0;
      bool live = false;
      if (r < c1) {
        const uint32_t d = row_doc(ix, r);
        live = !((ix.removed[d >> 5] >> (d & 31)) & 1u);
        if (live) {
#pragma unroll
          for (int f = 0; f < F; ++f) tf_sum += row_col(ix, r, f);
        }
      }
      if (ix.term_row_begin[t + 1] >= rend) {          // the 32 rows belong to term t (warp-uniform test)
        if (live) { s += tf_sum; ++n; }
      } else {
        // term boundaries inside this group: flush what belongs to t, then every lane finds its own row's term
        s = warp_sum_u64(s);
        uint32_t nn = n;
#pragma unroll
        for (int of = 16; of > 0; of >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, of);
        if (lane == 0 && nn) { atomicAdd(&df_live[t], s); atomicAdd(&live_rows[t], nn); }
        s = 0; n = 0;
        uint32_t tt = t;
        if (r < c1) {
          while (ix.term_row_begin[tt + 1] <= r) ++tt;
          if (live) { atomicAdd(&df_live[tt], (unsigned long long)tf_sum); atomicAdd(&live_rows[tt], 1u); }
        }
        // continue with the term of the group's last row
        uint32_t tl = __shfl_sync(0xffffffffu, tt, (int)(rend - r0 - 1));
        t = tl;
      }
    }
    s = warp_sum_u64(s);
#pragma unroll
    for (int of = 16; of > 0; of >>= 1) n += __shfl_xor_sync(0xffffffffu, n, of);
    if (lane == 0 && n) { atomicAdd(&df_live[t], s); atomicAdd(&live_rows[t], n); }
  }
}

// ------------------------------------------------------------------------------------------
// Warp iteration over a launch's virtual tile space
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t seg_of_tile(const uint64_t* __restrict__ tile_off, uint32_t sb, uint32_t se, uint64_t t) {
  // largest s in [sb, se) with tile_off[s] <= t  (segments with zero tiles are skipped)
  uint32_t a = sb, b = se;
  while (a < b) {
    uint32_t mid = (a + b) >> 1;
    if (tile_off[mid + 1] <= t) a = mid + 1; else b = mid;
  }
  return a;
}

// Marking pass of a side-path round.
// exact scheme (MODE_MULTI), words of the query = [summary: 1 bit per 1024 docs][bits A: 1 bit per doc]
//   [bits B: 1 bit per doc]: every list sets A (= "seen"); a doc seen again sets B (= "multi") and the
//   summary.  clear = 1 undoes the marks afterwards.
// primary scheme (MODE_SECONDARY rows), words of the query = row mask of the PRIMARY list, 4 words per
//   tile: each secondary doc is searched in the primary list (docs ascend inside a list) and, when
//   present, sets the bit of that primary row.  The scoring pass clears the words it consumes.
// Both count, per doc-range bin, the records the scoring pass will write (the row itself, + 1 for the
// doc's primary / first row, counted once).
template <int F>
__global__ void __launch_bounds__(CTA_THREADS) mark_kernel(const __grid_constant__ ScoreParams P, int clear) {
  const int lane = threadIdx.x & 31;
  const uint64_t w = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t W = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  // P.tile_off is the marking pass's own tile space here (primary lists own no tiles in it)
  const uint64_t tb = P.tile_off[P.seg_begin], te = P.tile_off[P.seg_end];
  const uint64_t T = te - tb;
  uint64_t t = tb + (T * w) / W;
  const uint64_t t1 = tb + (T * (w + 1)) / W;
  if (t >= t1) return;
  uint32_t s = seg_of_tile(P.tile_off, P.seg_begin, P.seg_end, t);
  unsigned long long xcount = 0, xtiles = 0;
  while (t < t1) {
    const Seg sg = P.segs[s];
    const uint64_t st0 = P.tile_off[s], st1 = P.tile_off[s + 1];
    const uint64_t tend = min(t1, st1);
    const bool exact = sg.mode == MODE_MULTI;
    if (tend > t && (exact || (sg.mode == MODE_SECONDARY && !clear))) {
      uint32_t* sm = P.bitmap + (size_t)(P.q_bmoff[sg.q] - P.round_bm0);
      uint32_t* ba = sm + P.bitmap_sum_words;
      uint32_t* bb = ba + P.bitmap_doc_words;
      const uint32_t bin_base = (uint32_t)(P.q_binoff[sg.q] - P.round_bin0);
      const uint32_t shift = P.q_shift[sg.q];
      const uint64_t abs0 = sg.row_begin / TILE_ROWS;
      const uint64_t rend = sg.row_begin + sg.n_rows;
      uint64_t pbeg = 0, pend = 0, pabs0 = 0;
      const uint2* pdir = nullptr;             // rank directory of the primary list, when it is a dense one
      if (!exact) {
        const Seg pr = P.segs[0xFFFFFFFFu - (uint32_t)(P.q_prim[sg.q] & 0xFFFFFFFFull)];
        pbeg = pr.row_begin; pend = pbeg + pr.n_rows; pabs0 = pbeg / TILE_ROWS;
        const uint32_t di = P.ix.term_dir[pr.term];
        if (di) pdir = P.ix.dir + (size_t)(di - 1) * P.ix.dir_words;
      }
      if (exact) xtiles += tend - t;
      for (; t < tend; ++t) {
        uint64_t row0 = (abs0 + (t - st0)) * TILE_ROWS + lane * 4;
        uint4 d = ldg_stream(tile_ptr(P.ix, abs0 + (t - st0)) + lane * 4);
        uint32_t dv[4] = {d.x, d.y, d.z, d.w};
        bool inr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) inr[j] = row0 + j >= sg.row_begin && row0 + j < rend;
        if (exact) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (!inr[j]) continue;
            const uint32_t doc = dv[j], bit = 1u << (doc & 31);
            if (clear) {
              ba[doc >> 5] = 0u; sm[doc >> 15] = 0u; bb[doc >> 5] = 0u;
            } else if (atomicOr(&ba[doc >> 5], bit) & bit) {      // seen before: a multi-event doc
              const uint32_t old = atomicOr(&bb[doc >> 5], bit);
              atomicOr(&sm[doc >> 15], 1u << ((doc >> 10) & 31));
              const uint32_t inc = (old & bit) ? 1u : 2u;
              atomicAdd(&P.bin_count[bin_base + (doc >> shift)], inc);
              xcount += inc;
            }
          }
        } else {
          // position of the lane's docs in the primary list: one directory load each for a dense list,
          // otherwise four lower-bound searches in lockstep
          uint64_t lo[4], hi[4];
          bool found[4] = {false, false, false, false};
          if (pdir) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              lo[j] = pend;
              if (inr[j]) {
                const uint2 e = __ldg(&pdir[dv[j] >> 5]);
                const uint32_t bit = 1u << (dv[j] & 31);
                found[j] = (e.y & bit) != 0u;
                lo[j] = pbeg + e.x + __popc(e.y & (bit - 1u));
              }
              hi[j] = lo[j];
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) { lo[j] = pbeg; hi[j] = inr[j] ? pend : pbeg; }
          }
          while (!pdir) {
            bool more = false;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (lo[j] < hi[j]) {
                const uint64_t mid = (lo[j] + hi[j]) >> 1;
                if (row_doc(P.ix, mid) < dv[j]) lo[j] = mid + 1; else hi[j] = mid;
                more = true;
              }
            }
            if (!more) break;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (!inr[j]) continue;
            uint32_t inc = 1u;
            const uint64_t r = lo[j];
            if (pdir ? found[j] : (r < pend && row_doc(P.ix, r) == dv[j])) {
              const uint32_t bit = 1u << (r & 31);
              const uint32_t old = atomicOr(&sm[(r / TILE_ROWS - pabs0) * 4 + ((r % TILE_ROWS) >> 5)], bit);
              inc = (old & bit) ? 1u : 2u;
            }
            atomicAdd(&P.bin_count[bin_base + (dv[j] >> shift)], inc);
            xcount += inc;
          }
        }
      }
    }
    t = tend;
    ++s;
  }
  xcount = warp_sum_u64(xcount);     // xtiles is warp-uniform already
  if (lane == 0 && xcount) atomicAdd(P.xcount, xcount);     // = sum of all bin capacities of the round
  if (lane == 0 && xtiles && !clear) atomicAdd(P.xtiles, xtiles);
}

// ------------------------------------------------------------------------------------------
// THE scoring kernel.  Each warp owns a contiguous span of the launch's tile space; a tile is
// 128 consecutive posting rows (aligned), each lane loads 4 rows of every column with one
// 128-bit streaming load (fully coalesced 512 B per column per warp).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t u4c(const uint4& v, int j) { return j == 0 ? v.x : j == 1 ? v.y : j == 2 ? v.z : v.w; }
// A lane's 4 rows of a tile.  Wide layout: a uint4 per tf / field-length column.  Narrow layout: a
// uint2 per field holding four u16 codes (tf << fl_bits | fl).
template <int F, bool NARROW> struct TileRegs;
template <int F> struct TileRegs<F, false> {
  uint4 dq;
  uint4 tq[F], lq[F];
  uint32_t mw;      // GMODE primary list: the row-mask word holding this lane's 4 bits
  uint32_t dw;      // IndexView::row_dead word holding this lane's 4 bits (0 when no doc is removed)
  __device__ __forceinline__ uint32_t tf(const IndexView&, int f, int j) const { return u4c(tq[f], j); }
  __device__ __forceinline__ uint32_t fl(const IndexView&, int f, int j) const { return u4c(lq[f], j); }
  // index into field f's table of saturated tf (row stride flcap)
  __device__ __forceinline__ uint32_t tabidx(int f, int j, uint32_t flcap) const { return u4c(tq[f], j) * flcap + u4c(lq[f], j); }
};
template <int F> struct TileRegs<F, true> {
  uint4 dq;
  uint2 cq[F];
  uint32_t mw;
  uint32_t dw;
  __device__ __forceinline__ uint32_t code(int f, int j) const {       // one PRMT
    return __byte_perm(j < 2 ? cq[f].x : cq[f].y, 0u, (j & 1) ? 0x4432u : 0x4410u);
  }
  __device__ __forceinline__ uint32_t tf(const IndexView& ix, int f, int j) const { return code(f, j) >> ix.fl_bits[f]; }
  __device__ __forceinline__ uint32_t fl(const IndexView& ix, int f, int j) const { return code(f, j) & ((1u << ix.fl_bits[f]) - 1u); }
  __device__ __forceinline__ uint32_t tabidx(int f, int j, uint32_t) const { return code(f, j); }   // flcap == 1 << fl_bits
};

// BM25::score (bm25.rs:60-93) for the 4 rows of a lane:  score += ((tf' * idf) * boost[x]) * eb.
// TABFULL: every (tf, fl) of the index is inside the shared-memory table of tf' (no range check,
// no division in the loop).  SIMPLE: all boosts and the expansion boost are exactly 1.0, so the
// two multiplications by 1.0 (exact identities) are skipped.
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
template <int F, bool TABFULL, int SIMPLE, bool NARROW>
__device__ __forceinline__ void bm25_rows(const ScoreParams& P, const uint32_t tbase, const TileRegs<F, NARROW>& R,
                                          double idf, double ebst, double (&sc)[4]) {
#pragma unroll
  for (int f = 0; f < F; ++f) {
    const uint32_t fbase = tbase + P.tab_boff[f];     // this lane's copy of field f's table (shared address)
    const uint32_t flcap = P.tab_flcap[f];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double tfn;
      if (TABFULL) {
        tfn = lds_f64(fbase + R.tabidx(f, j, flcap) * P.tab_stride);
      } else {
        const uint32_t tf = R.tf(P.ix, f, j), fl = R.fl(P.ix, f, j);
        tfn = (tf < P.tab_tfcap[f] && fl < flcap) ? lds_f64(fbase + (tf * flcap + fl) * P.tab_stride) : bm25_tf_slow(P, tf, fl, f);
      }
      double c = __dmul_rn(tfn, idf);
      if (SIMPLE == 0) c = __dmul_rn(c, P.boost[f]);     // x * 1.0 == x exactly, so levels 1 / 2 may skip
      if (SIMPLE < 2) c = __dmul_rn(c, ebst);
      if (TABFULL) {
        // table rows for tf = 0 hold +0.0 and idf / boosts / eb are finite, so c = +-0.0 there and
        // adding it is an exact identity: no select needed (the reference skips tf = 0, bm25.rs:73)
        if (f == 0) sc[j] = c; else sc[j] = __dadd_rn(sc[j], c);
      } else {
        const bool hit = R.tf(P.ix, f, j) > 0;
        if (f == 0) sc[j] = hit ? c : 0.0;
        else if (hit) sc[j] = __dadd_rn(sc[j], c);
      }
    }
  }
}

// Single-event ZeroToOne (score() + finalize(), zero_to_one.rs:44-126): max over fields of
// min(s/tf, 1) * tf / max(field_length, query_terms_len), floored at 0.
template <int F, bool NARROW>
__device__ __forceinline__ void z2o_rows(const ScoreParams& P, const TileRegs<F, NARROW>& R, double zs, uint32_t qtl,
                                         double (&sc)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) sc[j] = 0.0;
#pragma unroll
  for (int f = 0; f < F; ++f) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t tf = R.tf(P.ix, f, j), fl = R.fl(P.ix, f, j);
      if (tf > 0) sc[j] = fmax(z2o_entry(P.ix, zs, tf, fl, qtl), sc[j]);
    }
  }
}

// ZeroToOne side-path payload: the row's per-field (tf, fl) packed into 63 bits; bit 63 = escape
// (a value does not fit: the fold recovers the row by binary search and gathers it instead).
template <int F> struct Z2oPack {
  static constexpr int TFB = F == 1 ? 20 : F == 2 ? 11 : F == 3 ? 7 : 5;
  static constexpr int FLB = F == 1 ? 32 : F == 2 ? 20 : F == 3 ? 14 : 10;
  static_assert((TFB + FLB) * F <= 63, "payload layout");
};
template <int F>
__device__ __forceinline__ unsigned long long z2o_pack(const uint32_t (&tf)[F], const uint32_t (&fl)[F]) {
  unsigned long long p = 0;
  bool esc = false;
#pragma unroll
  for (int f = 0; f < F; ++f) {
    esc |= (Z2oPack<F>::TFB < 32 && (tf[f] >> Z2oPack<F>::TFB)) || (Z2oPack<F>::FLB < 32 && (fl[f] >> Z2oPack<F>::FLB));
    p |= ((unsigned long long)tf[f] | ((unsigned long long)fl[f] << Z2oPack<F>::TFB)) << (f * (Z2oPack<F>::TFB + Z2oPack<F>::FLB));
  }
  return esc ? (1ull << 63) : p;
}
template <int F>
__device__ __forceinline__ void z2o_unpack(unsigned long long p, uint32_t (&tf)[F], uint32_t (&fl)[F]) {
#pragma unroll
  for (int f = 0; f < F; ++f) {
    unsigned long long v = p >> (f * (Z2oPack<F>::TFB + Z2oPack<F>::FLB));
    tf[f] = (uint32_t)(v & ((1ull << Z2oPack<F>::TFB) - 1ull));
    fl[f] = (uint32_t)((v >> Z2oPack<F>::TFB) & ((1ull << Z2oPack<F>::FLB) - 1ull));
  }
}

// Per-segment state of the scoring loop.
struct SegCtx {
  uint64_t rbeg, rend;      // absolute row range of the segment
  double idf, ebst, zs;     // bm25.rs:35-58 / zero_to_one.rs:72
  uint32_t qtl;             // query_terms_len (query.rs:32)
  uint32_t seg;             // segment index (event order key of the side path)
  uint32_t slot;
  uint32_t mode;
  const uint32_t* sum;      // GMODE exact scheme: the query's summary bits / "multi" doc bits
  const uint32_t* bm;
  uint32_t* rowmask;        // GMODE primary scheme: row mask word of the segment's first tile
  uint32_t bin_base;        // GMODE: first bin of the query, relative to the round
  uint32_t shift;           // GMODE: log2(bin width in docs)
};

// One tile = 128 aligned rows; lane l owns rows 4l..4l+3 (one 128-bit load per u32 column, one
// 32-bit load per u8 column).
//   EDGE : the tile is shared with neighbouring segments -> per-row range mask
//   FAST : no full-result capture, complete BM25 table -> no per-row table range checks
template <int F, bool NARROW> struct TileGeom {
  static constexpr uint32_t WORDS = NARROW ? (TILE_ROWS + F * TILE_ROWS / 2) : (1 + 2 * F) * TILE_ROWS;   // u32 words per tile
};
__device__ __forceinline__ uint2 ldg_stream_u64(const uint32_t* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t rel_tile(const SegCtx& C, uint64_t tile_row) {
  return (uint32_t)(tile_row / TILE_ROWS - C.rbeg / TILE_ROWS);
}
__device__ __forceinline__ uint32_t* rowmask_word(const SegCtx& C, uint32_t ti, int lane) {
  return C.rowmask + (size_t)ti * 4 + (lane >> 3);
}
template <int F>
__device__ __forceinline__ void load_cols(TileRegs<F, false>& R, const uint32_t* base, int lane) {
#pragma unroll
  for (int f = 0; f < F; ++f) {
    R.tq[f] = ldg_stream(base + (1 + f) * TILE_ROWS + lane * 4);
    R.lq[f] = ldg_stream(base + (1 + F + f) * TILE_ROWS + lane * 4);
  }
}
template <int F>
__device__ __forceinline__ void load_cols(TileRegs<F, true>& R, const uint32_t* base, int lane) {
  // u16 code columns: 256 bytes each behind the doc column; this lane's 4 rows are one 8-byte load
#pragma unroll
  for (int f = 0; f < F; ++f) R.cq[f] = ldg_stream_u64(base + TILE_ROWS + f * (TILE_ROWS / 2) + lane * 2);
}
// `base` = first word of the tile's block, `maskp` = this lane's row-mask word of the tile (primary lists).
template <int F, bool NARROW>
__device__ __forceinline__ const uint32_t* tile_base(const ScoreParams& P, uint64_t tile_row) {
  return P.ix.post_blocks + (tile_row / TILE_ROWS) * (uint64_t)TileGeom<F, NARROW>::WORDS;
}
__device__ __forceinline__ uint32_t ld_dead(const uint32_t* p) {
  uint32_t v = 0;
  if (p != nullptr) asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ const uint32_t* dead_word(const IndexView& ix, uint64_t tile_row, int lane) {
  return ix.row_dead != nullptr ? ix.row_dead + (tile_row / TILE_ROWS) * 4 + (lane >> 3) : nullptr;
}
template <int F, bool GMODE, bool NARROW>
__device__ __forceinline__ void load_tile(const SegCtx& C, const uint32_t* base, const uint32_t* maskp, int lane, TileRegs<F, NARROW>& R,
                                          const uint32_t* deadp = nullptr) {
  // one contiguous block per tile: a single base address, immediate column offsets
  R.mw = 0;
  R.dw = ld_dead(deadp);
  if (GMODE && maskp != nullptr) {
    // the address does not depend on posting data: the mask word travels together with the tile
    // (maskp is null when the tile summary says the tile diverts nothing)
    asm volatile("ld.global.u32 %0, [%1];" : "=r"(R.mw) : "l"(maskp));
  }
  R.dq = ldg_stream(base + lane * 4);
  load_cols(R, base, lane);
}

// DEAD: the tile's removed-row bits came with the tile (R.dw, IndexView::row_dead) - no probe of the doc bitmap
template <int F, int SCORER, bool GMODE, bool EDGE, bool FAST, int SIMPLE, bool NARROW, bool DEAD = false>
__device__ __forceinline__ void compute_tile(const ScoreParams& P, const uint32_t s_tab, const SegCtx& C,
                                             const TileRegs<F, NARROW>& R, uint64_t tile_row, uint32_t ti, int lane,
                                             WarpAcc& acc, uint32_t& st_div) {
  const uint4 dq = R.dq;
  const uint32_t dv[4] = {dq.x, dq.y, dq.z, dq.w};
  uint32_t valid = 0xFu;
  if (EDGE) {
    const uint32_t lo = tile_row < C.rbeg ? (uint32_t)(C.rbeg - tile_row) : 0u;
    const uint32_t hi = tile_row + TILE_ROWS > C.rend ? (uint32_t)(C.rend - tile_row) : (uint32_t)TILE_ROWS;
    valid = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t r = lane * 4 + j;
      valid |= (r >= lo && r < hi) ? (1u << j) : 0u;
    }
  }
  if (DEAD) {                           // removed-but-not-vacuumed docs are skipped (query.rs:65)
    valid &= ~((R.dw >> ((lane & 7) * 4)) & 0xFu);
  } else if (P.ix.has_removed) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (((valid >> j) & 1u) && ((__ldg(&P.ix.removed[dv[j] >> 5]) >> (dv[j] & 31)) & 1u)) valid &= ~(1u << j);
  }
  double sc[4];
  uint32_t some = valid;
  if (SCORER == 0) {
    if (FAST) bm25_rows<F, true, SIMPLE, NARROW>(P, s_tab, R, C.idf, C.ebst, sc);
    else if (P.tab_full) bm25_rows<F, true, 0, NARROW>(P, s_tab, R, C.idf, C.ebst, sc);
    else bm25_rows<F, false, 0, NARROW>(P, s_tab, R, C.idf, C.ebst, sc);
    // Some(score) only if score > 0 (bm25.rs:89-92).  A double whose high word, read as a signed
    // integer, lies in (0, 0x7FF00000) is a positive finite number: when that holds for the four
    // rows of the lane (two integer min/max) the per-row f64 compares are skipped.
    bool check = true;
    if (FAST) {
      const int h0 = __double2hiint(sc[0]), h1 = __double2hiint(sc[1]), h2 = __double2hiint(sc[2]), h3 = __double2hiint(sc[3]);
      check = !(min(min(h0, h1), min(h2, h3)) > 0 && max(max(h0, h1), max(h2, h3)) < 0x7FF00000);
    }
    if (check) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (!(sc[j] > 0.0)) some &= ~(1u << j);
    }
  } else {
    z2o_rows<F, NARROW>(P, R, C.zs, C.qtl, sc);
  }
  if (GMODE) {
    // rows that must take the ordered per-doc fold instead: every live row of a secondary list
    // (scored or not: a None still marks the doc visited, query.rs:87), and the rows of the
    // primary list whose doc also occurs in a secondary list
    uint32_t dmask = valid;
    if (C.mode == MODE_PRIMARY) {
      dmask = (R.mw >> ((lane & 7) * 4)) & valid;
      // consumed: the pool is clean again for the next round (only the warp that scores the tile clears it)
      if (R.mw != 0u && (lane & 7) == 0) *rowmask_word(C, ti, lane) = 0u;
    } else if (C.mode != MODE_SECONDARY) {
      dmask = 0;
      bool maybe = true;
      if (!EDGE) {
        // docs ascend inside a list: the tile covers [first, last]; consult the 1024-doc summary bits
        const uint32_t b0 = __shfl_sync(0xffffffffu, dv[0], 0) >> 10, b1 = __shfl_sync(0xffffffffu, dv[3], 31) >> 10;
        if ((b0 >> 5) == (b1 >> 5)) {
          const uint32_t w = __ldg(&C.sum[b0 >> 5]);
          const uint32_t mask = (0xFFFFFFFFu >> (31u - (b1 & 31u))) & (0xFFFFFFFFu << (b0 & 31u));
          maybe = (w & mask) != 0u;
        }
      }
      if (maybe) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (((valid >> j) & 1u) && ((__ldg(&C.bm[dv[j] >> 5]) >> (dv[j] & 31)) & 1u)) dmask |= 1u << j;
      }
    }
    some &= ~dmask;
    if (dmask) {
      // one fat record per diverted row, written straight into the doc-range bin of the query
      // (overlapping the four cursor atomics of a lane was measured: no gain, more registers)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if ((dmask >> j) & 1u) {
          const uint32_t bin = C.bin_base + (dv[j] >> C.shift);
          const uint32_t pos = P.bin_off[bin] + atomicAdd(&P.bin_cursor[bin], 1u);
          unsigned long long pay;
          if (SCORER == 0) {
            pay = (unsigned long long)__double_as_longlong(sc[j]);     // the event's score (<= 0 / NaN = None)
          } else {
            uint32_t tfj[F], flj[F];
#pragma unroll
            for (int f = 0; f < F; ++f) { tfj[f] = R.tf(P.ix, f, j); flj[f] = R.fl(P.ix, f, j); }
            pay = z2o_pack<F>(tfj, flj);
          }
          if (pos < P.bin_off[bin + 1]) P.rec[pos] = make_uint4(dv[j], C.seg, (uint32_t)pay, (uint32_t)(pay >> 32));
          else atomicOr(P.out.error_flag, 2u);
        }
      }
      st_div += __popc(dmask);
    }
  }
  if (FAST) acc.template add4<false>(P.out, some, dv, sc, lane);
  else acc.template add4<true>(P.out, some, dv, sc, lane);
}

template <int F, int SCORER, bool GMODE, bool EDGE, bool FAST, int SIMPLE, bool NARROW>
__device__ __forceinline__ void process_tile(const ScoreParams& P, const uint32_t s_tab, const SegCtx& C,
                                             uint64_t tile_row, int lane, WarpAcc& acc, uint32_t& st_div) {
  TileRegs<F, NARROW> R;
  const uint32_t ti = (GMODE && C.mode == MODE_PRIMARY) ? rel_tile(C, tile_row) : 0u;
  uint32_t* maskp = (GMODE && C.mode == MODE_PRIMARY) ? rowmask_word(C, ti, lane) : nullptr;
  load_tile<F, GMODE, NARROW>(C, tile_base<F, NARROW>(P, tile_row), maskp, lane, R);
  compute_tile<F, SCORER, GMODE, EDGE, FAST, SIMPLE, NARROW>(P, s_tab, C, R, tile_row, ti, lane, acc, st_div);
}

// Interior tiles of one segment (every row belongs to the segment).  In the narrow layout a tile is
// only 5 + 2F registers per lane, so the loop is software-pipelined: the loads of tile i + 1 are in
// flight while tile i is scored (the spare tile behind the last row and the pad words of the row
// masks make the look-ahead load safe; nothing read ahead is used or cleared unless it is scored).
// The tile and mask addresses advance by constants: no 64-bit index arithmetic in the loop.
// (Round 3 tried a compile-time PRIMARY specialisation of this loop — no mask code at all for non-primary segments,
// no run-time flag for primary ones: class-G launch 18.4 ms instead of 15.6 in an alternating A/B on one box.  The
// shared loop below stays.)
template <int F, int SCORER, bool GMODE, bool FAST, int SIMPLE, bool NARROW, bool DEAD = false>
__device__ __forceinline__ void interior_tiles(const ScoreParams& P, const uint32_t s_tab, const SegCtx& C,
                                               uint64_t tile_row, uint32_t n_tiles, int lane, WarpAcc& acc,
                                               uint32_t& st_div) {
  if (NARROW) {
    if (n_tiles == 0) return;
    const uint32_t* base = tile_base<F, NARROW>(P, tile_row);
    const bool primary = GMODE && C.mode == MODE_PRIMARY;
    uint32_t ti = primary ? rel_tile(C, tile_row) : 0u;
    // Row-mask words of a primary list come from a multi-GB pool (DRAM latency): they are requested TWO
    // tiles ahead, the posting tiles (mostly L2 hits) one tile ahead.
    const uint32_t* maskp = primary ? rowmask_word(C, ti, lane) : nullptr;
    auto ld_mask = [&](const uint32_t* p) -> uint32_t {
      uint32_t v = 0;
      if (primary) asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
      return v;
    };
    // DEAD: the row_dead words travel with the tiles (the spare tiles behind the image cover the look-ahead)
    const uint32_t* deadp = DEAD ? dead_word(P.ix, tile_row, lane) : nullptr;
    TileRegs<F, NARROW> cur;
    load_tile<F, GMODE, NARROW>(C, base, nullptr, lane, cur, deadp);
    cur.mw = ld_mask(maskp);
    uint32_t mw1 = ld_mask(maskp + 4);               // mask of tile i + 1
#pragma unroll 2
    for (uint32_t i = 0; i < n_tiles; ++i) {
      TileRegs<F, NARROW> nxt;
      if (DEAD) deadp += 4;
      load_tile<F, GMODE, NARROW>(C, base + TileGeom<F, NARROW>::WORDS, nullptr, lane, nxt, deadp);
#if PB_L2_AHEAD
      // an image larger than L2 (cfg 3/4: 2.1 GB) streams from HBM: one tile of register look-ahead does
      // not cover DRAM latency, so the lines of the tile PB_L2_AHEAD tiles further on are pulled into L2
      if (P.l2_prefetch && i + PB_L2_AHEAD < n_tiles && lane < (int)(TileGeom<F, NARROW>::WORDS / 32))
        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + PB_L2_AHEAD * TileGeom<F, NARROW>::WORDS + lane * 32));
#endif
      const uint32_t mw2 = ld_mask(maskp + 8);       // mask of tile i + 2
      // a primary-list tile whose row-mask bits are all clear diverts nothing: score it exactly like a
      // single-list tile (the common case: ~93 % of the tiles of the cfg 1 batch)
      if (primary && !__any_sync(0xffffffffu, cur.mw != 0u))
        compute_tile<F, SCORER, false, false, FAST, SIMPLE, NARROW, DEAD>(P, s_tab, C, cur, 0, 0u, lane, acc, st_div);
      else
        compute_tile<F, SCORER, GMODE, false, FAST, SIMPLE, NARROW, DEAD>(P, s_tab, C, cur, 0, ti, lane, acc, st_div);
      cur = nxt;
      cur.mw = mw1;
      mw1 = mw2;
      base += TileGeom<F, NARROW>::WORDS;
      maskp += 4;
      ++ti;
    }
  } else {
    for (uint32_t i = 0; i < n_tiles; ++i, tile_row += TILE_ROWS)
      process_tile<F, SCORER, GMODE, false, FAST, SIMPLE, NARROW>(P, s_tab, C, tile_row, lane, acc, st_div);
  }
}

// Interior tiles of a single-list segment whose term is compact (IndexView::cpost): 2 + 2F bytes per row instead of
// 4 + 2F.  Same pipeline as above (tile i + 1 in flight while tile i is scored, L2 prefetch further ahead); the doc
// ordinals are rebuilt as base + u16 (two instructions per row) and the tile is scored by the same compute_tile.
template <int F> struct CompactRegs { uint2 d16; uint32_t base; uint32_t dw; uint2 cq[F]; };
template <int F>
__device__ __forceinline__ void load_compact(const uint32_t* cb, const uint32_t* bp, const uint32_t* deadp, int lane, CompactRegs<F>& R) {
  R.dw = ld_dead(deadp);
  R.d16 = ldg_stream_u64(cb + lane * 2);
#pragma unroll
  for (int f = 0; f < F; ++f) R.cq[f] = ldg_stream_u64(cb + (TILE_ROWS / 2) * (1 + f) + lane * 2);
  asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(R.base) : "l"(bp));      // one address per warp: a broadcast
}
template <int F, int SCORER, int SIMPLE>
__device__ __forceinline__ void interior_tiles_compact(const ScoreParams& P, const uint32_t s_tab, const SegCtx& C,
                                                       uint64_t tile_row, uint32_t n_tiles, int lane, WarpAcc& acc,
                                                       uint32_t& st_div) {
  if (n_tiles == 0) return;
  constexpr uint32_t CW = (TILE_ROWS / 2) * (1 + F);          // u32 words per compact tile
  const uint32_t* cb = P.ix.cpost + (tile_row / TILE_ROWS) * (uint64_t)CW;
  const uint32_t* bp = P.ix.cbase + tile_row / TILE_ROWS;
  const uint32_t* deadp = dead_word(P.ix, tile_row, lane);
  CompactRegs<F> cur;
  load_compact<F>(cb, bp, deadp, lane, cur);
#pragma unroll 2
  for (uint32_t i = 0; i < n_tiles; ++i) {
    CompactRegs<F> nxt;
    if (deadp != nullptr) deadp += 4;
    load_compact<F>(cb + CW, bp + 1, deadp, lane, nxt);        // the spare tiles behind the image make this safe
#if PB_L2_AHEAD
    if (P.l2_prefetch && i + PB_L2_AHEAD < n_tiles && lane < (int)(CW / 32))
      asm volatile("prefetch.global.L2 [%0];" ::"l"(cb + PB_L2_AHEAD * CW + lane * 32));
#endif
    TileRegs<F, true> R;
    R.mw = 0;
    R.dw = cur.dw;
    R.dq = make_uint4(cur.base + (cur.d16.x & 0xFFFFu), cur.base + (cur.d16.x >> 16),
                      cur.base + (cur.d16.y & 0xFFFFu), cur.base + (cur.d16.y >> 16));
#pragma unroll
    for (int f = 0; f < F; ++f) R.cq[f] = cur.cq[f];
    if (deadp != nullptr) compute_tile<F, SCORER, false, false, true, SIMPLE, true, true>(P, s_tab, C, R, 0, 0u, lane, acc, st_div);
    else compute_tile<F, SCORER, false, false, true, SIMPLE, true, false>(P, s_tab, C, R, 0, 0u, lane, acc, st_div);
    cur = nxt;
    cb += CW;
    ++bp;
  }
}

// CTA shape of the scoring kernel: no barrier after the table is loaded, so a CTA is just a bag of
// warps.  The host picks (threads per CTA, CTAs per SM) so that 24 warps are resident per SM
// whatever the shared-memory table costs: 3 x 256, 2 x 384 or 1 x 768 threads (80 registers per thread;
// 32 warps at 64 registers measured 5 % slower in an alternating A/B on one box: spills reach the loop).
#ifndef PB_SCORE_MAX_THREADS
#define PB_SCORE_MAX_THREADS 768
#endif
constexpr int SCORE_MAX_THREADS = PB_SCORE_MAX_THREADS;

template <int F, int SCORER, bool GMODE, bool NARROW>
__global__ void __launch_bounds__(SCORE_MAX_THREADS, 1) score_kernel(const __grid_constant__ ScoreParams P) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  if (SCORER == 0) {
    double* st = reinterpret_cast<double*>(s_raw);
    const uint32_t n = P.tab_total << P.tab_rep_shift;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) st[i] = P.tab[i >> P.tab_rep_shift];
  }
  const int lane = threadIdx.x & 31;
  // shared address of this lane's copy of table entry 0
  const uint32_t s_tab = smem_u32(s_raw) + ((uint32_t)lane & ((1u << P.tab_rep_shift) - 1u)) * 8u;
  __syncthreads();
  const uint64_t w = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t W = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint64_t T = P.tile_end - P.tile_begin;
  const uint64_t span0 = P.tile_begin + (T * w) / W;
  const uint64_t span1 = P.tile_begin + (T * (w + 1)) / W;
  if (span0 >= span1) return;
  uint64_t t = span0;
  uint32_t s = seg_of_tile(P.tile_off, P.seg_begin, P.seg_end, t);

  WarpAcc acc;
  acc.reset(NONE);
  bool acc_owned = false;
  uint32_t st_div = 0;
  // launch-wide: may the interior tiles take the check-free path?
  const bool fast = P.out.full_q == nullptr && (SCORER != 0 || P.tab_full);

  while (t < span1) {
    const Seg sg = P.segs[s];
    const uint64_t st0 = P.tile_off[s], st1 = P.tile_off[s + 1];
    const uint64_t tend = min(span1, st1);
    if (tend > t) {
      if (sg.q != acc.q) {
        if (acc.q != NONE) acc.flush(P.out, acc_owned, lane);
        acc.reset(sg.q);
        // class S: the query is exactly this segment; it is "owned" when its tiles are all ours
        acc_owned = !GMODE && st0 >= span0 && st1 <= span1;
      }
      SegCtx C;
      C.rbeg = sg.row_begin; C.rend = sg.row_begin + sg.n_rows;
      C.seg = s; C.slot = sg.slot; C.mode = sg.mode;
      C.idf = 0.0; C.ebst = 1.0; C.zs = 0.0; C.qtl = 0;
      C.sum = nullptr; C.bm = nullptr;
      const uint32_t explen = P.ix.term_byte_len[sg.term];
      bool simple = false;
      if (SCORER == 0) {
        C.idf = P.ix.term_idf[sg.term];
        C.ebst = P.ix.eb[explen - sg.qlen];
        simple = P.boosts_all_one && C.ebst == 1.0;
      } else {
        C.zs = z2o_term_score(explen, sg.qlen);
        C.qtl = (uint32_t)(P.query_term_off[sg.q + 1] - P.query_term_off[sg.q]);
      }
      C.bin_base = 0; C.shift = 0;
      C.rowmask = nullptr;
      if (GMODE) {
        C.rowmask = P.bitmap + (size_t)(P.q_bmoff[sg.q] - P.round_bm0);
        C.sum = C.rowmask;
        C.bm = C.sum + P.bitmap_sum_words + P.bitmap_doc_words;      // exact scheme: the "multi" plane
        C.bin_base = (uint32_t)(P.q_binoff[sg.q] - P.round_bin0);
        C.shift = P.q_shift[sg.q];
      }
      const uint64_t abs0 = C.rbeg / TILE_ROWS;
      // virtual tile range of the segment's fully covered (interior) tiles
      const uint64_t int0 = st0 + ((C.rbeg % TILE_ROWS) ? 1 : 0);
      const uint64_t int1 = st1 - ((C.rend % TILE_ROWS) ? 1 : 0);     // may be < int0 for a tiny segment
      const uint64_t ia = min(max(t, int0), tend), ib = max(min(tend, int1), ia);
      for (; t < ia; ++t)
        process_tile<F, SCORER, GMODE, true, false, 0, NARROW>(P, s_tab, C, (abs0 + (t - st0)) * TILE_ROWS, lane, acc, st_div);
      {
        const uint64_t row = (abs0 + (t - st0)) * TILE_ROWS;
        const uint32_t n = (uint32_t)(ib - t);
        if (!GMODE && NARROW && fast && P.ix.cpost != nullptr && P.ix.term_compact[sg.term]) {
          if constexpr (!GMODE && NARROW) {
            if (simple) interior_tiles_compact<F, SCORER, 2>(P, s_tab, C, row, n, lane, acc, st_div);
            else if (P.boosts_all_one) interior_tiles_compact<F, SCORER, 1>(P, s_tab, C, row, n, lane, acc, st_div);
            else interior_tiles_compact<F, SCORER, 0>(P, s_tab, C, row, n, lane, acc, st_div);
          }
        } else if (fast && NARROW && P.ix.row_dead != nullptr) {      // docs are removed: their row bits come with the tiles
          if (simple) interior_tiles<F, SCORER, GMODE, true, 2, NARROW, NARROW>(P, s_tab, C, row, n, lane, acc, st_div);
          else if (P.boosts_all_one) interior_tiles<F, SCORER, GMODE, true, 1, NARROW, NARROW>(P, s_tab, C, row, n, lane, acc, st_div);
          else interior_tiles<F, SCORER, GMODE, true, 0, NARROW, NARROW>(P, s_tab, C, row, n, lane, acc, st_div);
        } else if (fast) {
          if (simple) interior_tiles<F, SCORER, GMODE, true, 2, NARROW>(P, s_tab, C, row, n, lane, acc, st_div);
          else if (P.boosts_all_one) interior_tiles<F, SCORER, GMODE, true, 1, NARROW>(P, s_tab, C, row, n, lane, acc, st_div);
          else interior_tiles<F, SCORER, GMODE, true, 0, NARROW>(P, s_tab, C, row, n, lane, acc, st_div);
        } else {
          interior_tiles<F, SCORER, GMODE, false, 0, NARROW>(P, s_tab, C, row, n, lane, acc, st_div);
        }
        t = ib;
      }
      for (; t < tend; ++t)
        process_tile<F, SCORER, GMODE, true, false, 0, NARROW>(P, s_tab, C, (abs0 + (t - st0)) * TILE_ROWS, lane, acc, st_div);
    }
    ++s;
  }
  if (acc.q != NONE) acc.flush(P.out, acc_owned, lane);
  if (GMODE) {
    unsigned long long d = warp_sum_u64(st_div);
    if (lane == 0 && d) atomicAdd(&P.stats[ST_ROWS_DIVERTED], d);
  }
}

// ------------------------------------------------------------------------------------------
// Side path: fold the events of docs that were hit by more than one (query term, expansion).
// Records are sorted by (round slot, doc); the events of one doc are consecutive.  Within a
// doc they must be applied in (query_term_index, expansion rank) order = ascending segment
// index (segments are generated in that order), which is the high half of the record value.
// ------------------------------------------------------------------------------------------
struct FoldParams {
  ScoreParams S;
  const unsigned long long* key;   // sorted
  const unsigned long long* val;
  uint8_t* flags;                  // [n] scratch of the slow per-doc fold (docs with more than 32 events)
  uint32_t n;
};

template <int F>
__device__ __forceinline__ double bm25_row_score(const ScoreParams& P, const Seg& sg, uint32_t row) {
  const uint32_t explen = P.ix.term_byte_len[sg.term];
  const double idf = P.ix.term_idf[sg.term];
  const double ebst = P.ix.eb[explen - sg.qlen];
  double score = 0.0;
#pragma unroll
  for (int f = 0; f < F; ++f) {
    uint32_t tf = row_tf<F>(P.ix, row, f), fl = row_fl<F>(P.ix, row, f);
    if (tf > 0) {
      double tfn = (tf < P.tab_tfcap[f] && fl < P.tab_flcap[f]) ? P.tab[P.tab_off[f] + tf * P.tab_flcap[f] + fl]
                                                                 : bm25_tf_slow(P, tf, fl, f);
      double c = __dmul_rn(__dmul_rn(__dmul_rn(tfn, idf), P.boost[f]), ebst);
      score = __dadd_rn(score, c);
    }
  }
  return score;
}

// next event of the doc in ascending value order: smallest val > last (vals are distinct)
__device__ __forceinline__ bool next_event(const unsigned long long* val, uint32_t a, uint32_t b, bool first,
                                           unsigned long long last, unsigned long long* out) {
  bool found = false;
  unsigned long long best = 0;
  for (uint32_t j = a; j < b; ++j) {
    unsigned long long v = val[j];
    if ((first || v > last) && (!found || v < best)) { best = v; found = true; }
  }
  *out = best;
  return found;
}

// Slow per-thread fold of ONE doc's events [i, e) straight from global memory: used only for docs
// with more than 32 events (they do not fit a warp window).  Same rules as the warp path below.
template <int F, int SCORER>
__device__ __noinline__ bool fold_group_slow(const FoldParams& FP, uint32_t i, uint32_t e, uint32_t q, double* out) {
  const ScoreParams& P = FP.S;
  bool has = false;
  double result = 0.0;
  if (SCORER == 0) {
    uint32_t cur = NONE;
    unsigned long long last = 0, v;
    bool first_ev = true;
    while (next_event(FP.val, i, e, first_ev, last, &v)) {
      first_ev = false; last = v;
      const Seg sg = P.segs[(uint32_t)(v >> 32)];
      bool firstq = (sg.qti != cur);
      cur = sg.qti;
      double sc = bm25_row_score<F>(P, sg, (uint32_t)v);
      if (!(sc > 0.0)) continue;
      if (!has) { result = sc; has = true; }
      else if (firstq) result = __dadd_rn(result, sc);
      else result = fmax(result, sc);
    }
  } else {
    has = true;
    const uint32_t qtl = (uint32_t)(P.query_term_off[q + 1] - P.query_term_off[q]);
    const uint32_t ne = e - i;
    uint8_t* fg = FP.flags + i;            // per event: bit 0 = looked at, bit 1 = accepted (this doc's slice of the scratch)
#pragma unroll 1
    for (int x = 0; x < F; ++x) {
      double accx = 0.0;
      for (uint32_t j = 0; j < ne; ++j) fg[j] = 0;
      for (uint32_t step = 0; step < ne; ++step) {
        int bj = -1; double bs = 0.0; unsigned long long bv = 0; uint32_t btf = 0, bfl = 0, bterm = 0, bqti = 0;
        for (uint32_t j = 0; j < ne; ++j) {
          if (fg[j] & 1u) continue;
          unsigned long long v = FP.val[i + j];
          uint32_t row = (uint32_t)v;
          uint32_t tf = row_tf<F>(P.ix, row, x);
          if (tf == 0) { fg[j] |= 1u; continue; }
          const Seg sg = P.segs[(uint32_t)(v >> 32)];
          double sc = z2o_term_score(P.ix.term_byte_len[sg.term], sg.qlen);
          if (bj < 0 || sc > bs || (sc == bs && v < bv)) {
            bj = (int)j; bs = sc; bv = v; btf = tf; bfl = row_fl<F>(P.ix, row, x); bterm = sg.term; bqti = sg.qti;
          }
        }
        if (bj < 0) break;
        fg[bj] |= 1u;
        bool consumed = false; uint32_t used = 0;
        for (uint32_t j = 0; j < ne; ++j) {
          if (!(fg[j] & 2u)) continue;
          const Seg sg = P.segs[(uint32_t)(FP.val[i + j] >> 32)];
          consumed |= (sg.qti == bqti);
          used += (sg.term == bterm);
        }
        if (consumed || used >= btf) continue;
        fg[bj] |= 2u;
        accx = __dadd_rn(accx, z2o_entry(P.ix, bs, btf, bfl, qtl));
      }
      result = fmax(accx, result);
    }
  }
  *out = result;
  return has;
}

__device__ __forceinline__ double shfl_f64(double v, int src) {
  return __longlong_as_double((long long)shfl_u64((uint64_t)__double_as_longlong(v), src));
}

// The fold of one window of <= 32 events (lane = event).  The events of a doc occupy consecutive
// lanes (`hm` = mask of the lanes that start a doc, `n_act` = lanes in use).  PRESORTED: the lanes
// of a doc are already in event order (segment order).
template <int F, int SCORER, bool PRESORTED>
__device__ __forceinline__ void fold_window(const ScoreParams& P, WarpAcc& acc, int lane, bool in, bool head, uint32_t hm,
                                            uint32_t advance, uint32_t q, uint32_t doc, unsigned long long val,
                                            double ev_score, uint32_t e_qti, uint32_t e_term, uint32_t (&tfv)[F],
                                            uint32_t (&flv)[F]) {
  // group geometry
  const int my_head = 31 - __clz(hm & (0xFFFFFFFFu >> (31 - lane)));
  const uint32_t above = hm & (0xFFFFFFFEu << my_head);
  const int g_end = above ? (__ffs(above) - 1) : (int)advance;
  const int size = in ? g_end - my_head : 0;
  const int pos = lane - my_head;
  const uint32_t gmask = in ? (((size >= 32) ? 0xFFFFFFFFu : ((1u << size) - 1u)) << my_head) : 0u;
  int maxsize = size;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) maxsize = max(maxsize, __shfl_xor_sync(0xffffffffu, maxsize, o));

  // rank of the event inside its doc in processing order
  int rank = pos;
  if (!(PRESORTED && SCORER == 0)) {
  rank = 0;
  for (int m = 0; m < maxsize; ++m) {
    const int src = min(my_head + m, 31);
    const unsigned long long ov = shfl_u64(val, src);
    bool before;
    if (SCORER == 0) {
      before = ov < val;
    } else {
      const double os = shfl_f64(ev_score, src);
      before = os > ev_score || (os == ev_score && ov < val);
    }
    if (m < size && m != pos && before) ++rank;
  }
  // pull the event that belongs at this lane's position
  int src = lane;
  for (int m = 0; m < maxsize; ++m) {
    const int c = min(my_head + m, 31);
    const int rc = __shfl_sync(0xffffffffu, rank, c);
    if (m < size && rc == pos) src = c;
  }
  ev_score = shfl_f64(ev_score, src);
  e_qti = __shfl_sync(0xffffffffu, e_qti, src);
  e_term = __shfl_sync(0xffffffffu, e_term, src);
#pragma unroll
  for (int f = 0; f < F; ++f) { tfv[f] = __shfl_sync(0xffffffffu, tfv[f], src); flv[f] = __shfl_sync(0xffffffffu, flv[f], src); }
  }

  bool has = false;
  double result = 0.0;
  if (SCORER == 0) {
    uint32_t cur = NONE;
    for (int m = 0; m < maxsize; ++m) {
      const int c = min(my_head + m, 31);
      const double sm = shfl_f64(ev_score, c);
      const uint32_t qm = __shfl_sync(0xffffffffu, e_qti, c);
      if (m < size) {
        const bool firstq = qm != cur;
        cur = qm;
        if (sm > 0.0) {                          // None otherwise: the doc is only marked visited
          if (!has) { result = sm; has = true; }
          else if (firstq) result = __dadd_rn(result, sm);
          else result = fmax(result, sm);
        }
      }
    }
  } else {
    has = in;
    const uint32_t qtl = in ? (uint32_t)(P.query_term_off[q + 1] - P.query_term_off[q]) : 0u;
#pragma unroll
    for (int x = 0; x < F; ++x) {
      double accx = 0.0;
      bool accepted = false;                     // this lane's entry was accepted for field x
      for (int m = 0; m < maxsize; ++m) {
        const int c = min(my_head + m, 31);
        const uint32_t c_qti = __shfl_sync(0xffffffffu, e_qti, c);
        const uint32_t c_term = __shfl_sync(0xffffffffu, e_term, c);
        const uint32_t c_tf = __shfl_sync(0xffffffffu, tfv[x], c);
        const uint32_t conflict = __ballot_sync(0xffffffffu, accepted && e_qti == c_qti) & gmask;
        const uint32_t used = __popc(__ballot_sync(0xffffffffu, accepted && e_term == c_term) & gmask);
        const bool ok = m < size && c_tf > 0 && conflict == 0u && used < c_tf;
        double contrib = 0.0;
        if (ok && pos == m) { accepted = true; contrib = z2o_entry(P.ix, ev_score, tfv[x], flv[x], qtl); }
        contrib = shfl_f64(contrib, c);
        if (ok) accx = __dadd_rn(accx, contrib);
      }
      result = fmax(accx, result);
    }
  }
  has = has && head;
  // emit: lanes may belong to different queries (sorted, so at most a few switches)
  uint32_t mres = __ballot_sync(0xffffffffu, has);
  while (mres) {
    int l = __ffs(mres) - 1;
    uint32_t ql = __shfl_sync(0xffffffffu, q, l);
    bool mine = has && q == ql;
    if (acc.q != ql) {
      if (acc.q != NONE) acc.flush(P.out, false, lane);
      acc.reset(ql);
    }
    acc.add(P.out, mine, doc, result, lane);
    mres &= ~__ballot_sync(0xffffffffu, mine);
  }
}

// Warp-cooperative fold.  A warp walks its span of the sorted records in windows of 32: lane =
// record.  Every lane gathers its own event (segment descriptor, posting row) — 32 independent
// gathers in flight — then the events of a doc (consecutive lanes) are put in processing order
// with a few shuffles and combined:
//   BM25     order = segment index = (query term, expansion rank): has/first/max fold of
//            query.rs:150-164, None events still marking the doc visited.
//   ZeroToOne order = (entry score desc, event order asc) = the stable sort of zero_to_one.rs:98;
//            one candidate per doc per step; the lanes holding accepted entries vote whether the
//            candidate's query term is consumed / its term's pool exhausted (zero_to_one.rs:101-115).
// A window always starts at a doc boundary; a doc cut by the window end is deferred to the next
// window; docs with more than 32 events take fold_group_slow.
template <int F, int SCORER>
__global__ void __launch_bounds__(CTA_THREADS) fold_kernel(const __grid_constant__ FoldParams FP) {
  const ScoreParams& P = FP.S;
  const int lane = threadIdx.x & 31;
  const uint64_t w = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t W = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint32_t n = FP.n;
  uint32_t base = (uint32_t)(((uint64_t)n * w) / W);
  const uint32_t stop = (uint32_t)(((uint64_t)n * (w + 1)) / W);     // docs whose head is < stop are ours
  // move to the first doc boundary at or after `base`
  while (base > 0 && base < n && FP.key[base] == FP.key[base - 1]) ++base;
  WarpAcc acc;
  acc.reset(NONE);
  const uint64_t doc_mask = (1ull << P.doc_bits) - 1ull;

  while (base < stop) {
    const uint32_t i = base + lane;
    const uint32_t n_in = min(32u, n - base);
    bool in = lane < (int)n_in;
    unsigned long long key = in ? FP.key[i] : ~0ull;
    const unsigned long long key_up = __shfl_up_sync(0xffffffffu, key, 1);
    bool head = in && (lane == 0 || key_up != key);
    uint32_t hm = __ballot_sync(0xffffffffu, head);
    // `lim` = first lane that is NOT processed in this window:
    //  - docs whose head is at or after `stop` belong to the next warp;
    //  - a doc cut by the window end is deferred to the next window (or, if it fills the whole
    //    window, i.e. has more than 32 events, folded by the slow path right here).
    uint32_t lim = 32;
    if (stop - base < 32u) {
      const uint32_t hb = hm & (0xFFFFFFFFu << (stop - base));
      if (hb) lim = (uint32_t)(__ffs(hb) - 1);
    }
    if (lim == 32 && n_in == 32 && base + 32 < n) {
      const unsigned long long knext = FP.key[base + 32];
      if (__shfl_sync(0xffffffffu, key, 31) == knext) {
        const int h31 = 31 - __clz(hm);               // head lane of the cut doc
        if (h31 == 0) {
          uint32_t e = base + 32;
          while (e < n && FP.key[e] == key) ++e;       // every lane holds the same key here
          double r = 0.0;
          bool hs = false;
          const uint32_t q0 = P.segs[(uint32_t)(FP.val[base] >> 32)].q;
          if (lane == 0) hs = fold_group_slow<F, SCORER>(FP, base, e, q0, &r);
          if (acc.q != q0) { if (acc.q != NONE) acc.flush(P.out, false, lane); acc.reset(q0); }
          acc.add(P.out, lane == 0 && hs, (uint32_t)(key & doc_mask), r, lane);
          base = e;
          continue;
        }
        lim = (uint32_t)h31;
      }
    }
    if (lim < 32) {
      in = in && lane < (int)lim;
      head = head && lane < (int)lim;
      hm &= (1u << lim) - 1u;
    }
    const uint32_t advance = min(lim, n_in);
    // every lane gathers its own event
    unsigned long long val = in ? FP.val[i] : 0ull;
    Seg sg;
    sg.q = NONE; sg.qti = 0; sg.term = 0; sg.qlen = 0;
    if (in) sg = P.segs[(uint32_t)(val >> 32)];
    const uint32_t row = (uint32_t)val;
    const uint32_t doc = (uint32_t)(key & doc_mask);
    uint32_t q = sg.q;
    double ev_score = 0.0;      // BM25: the event's score; ZeroToOne: the entry score of the term
    uint32_t tfv[F], flv[F];
#pragma unroll
    for (int f = 0; f < F; ++f) { tfv[f] = 0; flv[f] = 0; }
    if (in) {
      if (SCORER == 0) {
        ev_score = bm25_row_score<F>(P, sg, row);
      } else {
        ev_score = z2o_term_score(P.ix.term_byte_len[sg.term], sg.qlen);
#pragma unroll
        for (int f = 0; f < F; ++f) { tfv[f] = row_tf<F>(P.ix, row, f); flv[f] = row_fl<F>(P.ix, row, f); }
      }
    }
    fold_window<F, SCORER, false>(P, acc, lane, in, head, hm, advance, q, doc, val, ev_score, (uint32_t)sg.qti, sg.term, tfv, flv);
    base += advance;
  }
  if (acc.q != NONE) acc.flush(P.out, false, lane);
}

// Row of `doc` inside a segment's posting list (docs ascend): used only on fallback paths.
template <int F>
__device__ __forceinline__ uint32_t find_row(const ScoreParams& P, const Seg& sg, uint32_t doc) {
  uint64_t lo = sg.row_begin, hi = sg.row_begin + sg.n_rows;
  while (lo < hi) {
    uint64_t mid = (lo + hi) >> 1;
    if (row_doc(P.ix, mid) < doc) lo = mid + 1; else hi = mid;
  }
  return (uint32_t)lo;
}

// ------------------------------------------------------------------------------------------
// Bin fold: the side path without a global sort.  The scoring pass has written the diverted
// events as fat records into per-query doc-range bins sized for ~8 records.  A warp packs as
// many consecutive whole bins as fit into its 32 lanes (lane = record), orders the window by
// (bin, doc, event order) with a 15-step bitonic network of shuffles, and folds every doc with
// fold_window.  Nothing is gathered from the posting columns: BM25 records carry the score,
// ZeroToOne records carry the packed (tf, fl).  A bin holding more than 32 records (doc ids
// clustered far beyond the uniform expectation) is handed to the legacy sorted path.
// ------------------------------------------------------------------------------------------
template <int F, int SCORER>
__global__ void __launch_bounds__(CTA_THREADS) binfold_kernel(const __grid_constant__ ScoreParams P) {
  const int lane = threadIdx.x & 31;
  const uint64_t w = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t W = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  uint32_t b = (uint32_t)(((uint64_t)P.n_bins * w) / W);
  const uint32_t b1 = (uint32_t)(((uint64_t)P.n_bins * (w + 1)) / W);
  WarpAcc acc;
  acc.reset(NONE);
  while (b < b1) {
    // lane l looks at bin b + l
    const uint32_t bl = b + lane;
    const uint32_t cnt = bl < b1 ? P.bin_cursor[bl] : 64u;
    const uint32_t boff = bl < b1 ? P.bin_off[bl] : 0u;
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const uint32_t take = __ballot_sync(0xffffffffu, incl <= PB_WINDOW_MAX && bl < b1);   // a prefix of the lanes
    const int nb_take = __popc(take);
    if (nb_take == 0) {
      // bin b alone overflows a window: hand its records to the legacy sorted path
      const uint32_t c0 = __shfl_sync(0xffffffffu, cnt, 0), o0 = __shfl_sync(0xffffffffu, boff, 0);
      for (uint32_t i = lane; i < c0; i += 32) {
        const uint4 r = P.rec[o0 + i];
        const Seg sg = P.segs[r.y];
        const uint32_t row = find_row<F>(P, sg, r.x);
        const uint32_t pos = atomicAdd(P.rec_count, 1u);
        if (pos < P.rec_cap) {
          P.rec_key[pos] = ((unsigned long long)sg.slot << P.doc_bits) | r.x;
          P.rec_val[pos] = ((unsigned long long)r.y << 32) | row;
        } else {
          atomicOr(P.out.error_flag, 2u);
        }
      }
      b += 1;
      continue;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, nb_take - 1);
    if (total == 0) { b += nb_take; continue; }
    // record lane i -> (bin t, index inside the bin)
    const bool in = lane < (int)total;
    int lo = 0, hi = nb_take;             // smallest t with incl[t] > lane
#pragma unroll
    for (int it = 0; it < 6; ++it) {             // a range of up to 32 bins needs 6 halvings
      const int mid = (lo + hi) >> 1;
      const uint32_t v = __shfl_sync(0xffffffffu, incl, min(mid, 31));
      if (lo < hi) { if (v <= (uint32_t)lane) lo = mid + 1; else hi = mid; }
    }
    const int t = min(lo, nb_take - 1);
    const uint32_t t_incl = __shfl_sync(0xffffffffu, incl, t), t_cnt = __shfl_sync(0xffffffffu, cnt, t);
    const uint32_t t_off = __shfl_sync(0xffffffffu, boff, t);
    uint4 r = make_uint4(0, 0, 0, 0);
    Seg sg;
    sg.q = NONE; sg.qti = 0; sg.term = 0; sg.qlen = 0; sg.row_begin = 0; sg.n_rows = 0; sg.slot = 0;
    unsigned long long key = ~0ull;
    if (in) {
      r = P.rec[t_off + (lane - (t_incl - t_cnt))];
      sg = P.segs[r.y];
      const unsigned long long segrel = (unsigned long long)r.y - P.q_gsegoff[sg.q];
      if (segrel >> 27) atomicOr(P.out.error_flag, 8u);       // > 2^27 posting lists in one query
      key = ((unsigned long long)t << 59) | ((unsigned long long)r.x << 27) | (segrel & 0x7FFFFFFull);
    }
    // bitonic sort of the window by key; `src` remembers where each record came from
    int src = lane;
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) {
        const unsigned long long okey = shfl_u64(key, lane ^ j);
        const int osrc = __shfl_xor_sync(0xffffffffu, src, j);
        const bool asc = (lane & k) == 0, lower = (lane & j) == 0;
        const bool other_less = okey < key || (okey == key && osrc < src);
        if ((lower == asc) == other_less) { key = okey; src = osrc; }
      }
    }
    // bring the payload along
    const uint32_t s_seg = __shfl_sync(0xffffffffu, r.y, src);
    const unsigned long long pay = shfl_u64(((unsigned long long)r.w << 32) | r.z, src);
    const uint32_t q = __shfl_sync(0xffffffffu, sg.q, src);
    const uint32_t e_qti = __shfl_sync(0xffffffffu, (uint32_t)sg.qti, src);
    const uint32_t e_term = __shfl_sync(0xffffffffu, sg.term, src);
    const uint32_t e_qlen = __shfl_sync(0xffffffffu, sg.qlen, src);
    const uint32_t doc = (uint32_t)(key >> 27);
    const unsigned long long gk = key >> 27;                      // (bin, doc)
    const unsigned long long gk_up = __shfl_up_sync(0xffffffffu, gk, 1);
    const bool head = in && (lane == 0 || gk_up != gk);
    const uint32_t hm = __ballot_sync(0xffffffffu, head);
    double ev_score = 0.0;
    uint32_t tfv[F], flv[F];
#pragma unroll
    for (int f = 0; f < F; ++f) { tfv[f] = 0; flv[f] = 0; }
    if (in) {
      if (SCORER == 0) {
        ev_score = __longlong_as_double((long long)pay);
      } else {
        ev_score = z2o_term_score(P.ix.term_byte_len[e_term], e_qlen);
        if (pay >> 63) {                                           // escape: gather the row
          const Seg sg2 = P.segs[s_seg];
          const uint32_t row = find_row<F>(P, sg2, doc);
#pragma unroll
          for (int f = 0; f < F; ++f) { tfv[f] = row_tf<F>(P.ix, row, f); flv[f] = row_fl<F>(P.ix, row, f); }
        } else {
          z2o_unpack<F>(pay, tfv, flv);
        }
      }
    }
    fold_window<F, SCORER, true>(P, acc, lane, in, head, hm, total, q, doc, (unsigned long long)lane, ev_score, e_qti, e_term,
                                 tfv, flv);
    b += nb_take;
  }
  if (acc.q != NONE) acc.flush(P.out, false, lane);
}

}  // namespace pbk
