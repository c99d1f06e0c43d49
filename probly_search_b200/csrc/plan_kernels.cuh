// Non-template kernels of the query path: trie descent, rank-directory build and the planning stage
// (classification of queries, segment descriptors, doc-range bins).  Included by engine.cu ONLY: the
// per-field-count kernels live in kernels.cuh and are compiled once per F in kernels_f.cu.
#pragma once
#include "kernels.cuh"
#include "union_kernels.cuh"

namespace pbk {
// ------------------------------------------------------------------------------------------
// Trie descent + prefix expansion
// ------------------------------------------------------------------------------------------
// One thread per query term.  find_inverted_index_node (index.rs:300-318) with a binary search
// over the node's char-sorted edges instead of the sibling-list scan (index.rs:321-337).
// Output: the DFS term range [lo, hi) = expand_term's result (query.rs:109-147), already in the
// reference's expansion order; lo == hi when nothing matches or the token is empty (query.rs:35).
__global__ void descend_kernel(IndexView ix, const uint8_t* __restrict__ term_bytes,
                               const uint64_t* __restrict__ term_byte_off, uint64_t n_qterms,
                               uint32_t* __restrict__ qt_lo, uint32_t* __restrict__ qt_hi,
                               uint32_t* __restrict__ qt_len) {
  uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (t >= n_qterms) return;
  const uint8_t* p = term_bytes + term_byte_off[t];
  const uint8_t* e = term_bytes + term_byte_off[t + 1];
  uint32_t len = (uint32_t)(e - p);
  qt_len[t] = len;
  uint32_t node = 0;
  bool ok = len > 0;
  while (ok && p < e) {
    uint32_t c = *p++;
    if (c >= 0x80) {                       // UTF-8 (validated on the host) -> Unicode scalar
      int extra = (c >= 0xF0) ? 3 : (c >= 0xE0) ? 2 : 1;
      c &= (0x3Fu >> extra);
      for (int i = 0; i < extra && p < e; ++i) c = (c << 6) | (*p++ & 0x3Fu);
    }
    uint32_t lo = ix.node_edge_begin[node], hi = ix.node_edge_begin[node + 1];
    while (lo < hi) {
      uint32_t mid = (lo + hi) >> 1;
      if (ix.edge_char[mid] < c) lo = mid + 1; else hi = mid;
    }
    if (lo < ix.node_edge_begin[node + 1] && ix.edge_char[lo] == c) node = ix.edge_child[lo];
    else ok = false;
  }
  qt_lo[t] = ok ? ix.node_term_lo[node] : 0u;
  qt_hi[t] = ok ? ix.node_term_hi[node] : 0u;
}

// Rank directory build (see IndexView::dir).  Pass 1: one warp-strided walk over the rows of every
// dense term sets the doc bits.  Pass 2: one CTA per dense term turns the per-word popcounts into
// exclusive prefix sums.
__global__ void dir_bits_kernel(IndexView ix, const uint32_t* __restrict__ dense_terms, uint32_t n_dense, uint2* __restrict__ dir) {
  for (uint32_t i = blockIdx.y; i < n_dense; i += gridDim.y) {
    const uint32_t t = dense_terms[i];
    const uint64_t a = ix.term_row_begin[t], b = ix.term_row_begin[t + 1];
    uint2* d = dir + (size_t)i * ix.dir_words;
    for (uint64_t r = a + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < b; r += (uint64_t)gridDim.x * blockDim.x) {
      const uint32_t doc = row_doc(ix, r);
      atomicOr(&d[doc >> 5].y, 1u << (doc & 31));
    }
  }
}
__global__ void __launch_bounds__(1024) dir_rank_kernel(uint32_t dir_words, uint2* __restrict__ dir) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_base;
  uint2* d = dir + (size_t)blockIdx.x * dir_words;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (uint32_t w0 = 0; w0 < dir_words; w0 += blockDim.x) {
    const uint32_t w = w0 + threadIdx.x;
    const uint32_t c = w < dir_words ? __popc(d[w].y) : 0u;
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += v; }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t v = s_warp[threadIdx.x], x = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, x, o); if (threadIdx.x >= o) x += u; }
      s_warp[threadIdx.x] = x - v;              // exclusive prefix of the warp totals
    }
    __syncthreads();
    const uint32_t base = s_base;
    if (w < dir_words) d[w].x = base + s_warp[threadIdx.x >> 5] + incl - c;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_base = base + s_warp[threadIdx.x >> 5] + incl;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// Planning
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t first_live_term(const IndexView& ix, uint32_t lo, uint32_t hi) {
  // smallest t in [lo, hi) with df_live > 0, i.e. live_prefix[t+1] > live_prefix[lo]
  uint32_t base = ix.live_prefix[lo];
  uint32_t a = lo, b = hi;
  while (a < b) {
    uint32_t mid = (a + b) >> 1;
    if (ix.live_prefix[mid + 1] > base) b = mid; else a = mid + 1;
  }
  return a;
}

// What plan_query_kernel needs to route a query to the dense union path.
struct UPlan {
  uint32_t enabled;
  uint32_t allow_gen;                          // the two-candidate kernel fits shared memory at this field count
  uint32_t max_term_bytes, max_tf;             // of the index: longest term, largest tf over all fields
  unsigned long long min_rows;                 // a class-U query streams at least this many rows
  UQuery* uq;                                  // [n_queries]
  unsigned long long* q_isu;                   // [n_queries + 1] class U without overlapping term ranges
  unsigned long long* q_isu2;                  // [n_queries + 1] class U with overlaps (general kernel)
  const unsigned long long* liverowcnt_prefix; // [n_terms + 1] rows whose doc is live, before term t
  const unsigned long long* dflive_prefix;     // [n_terms + 1] live occurrence counts before term t
  unsigned long long* stats;                   // [ST_COUNT] of the union launch
};

// One thread per query.  Classifies the query by the number of live posting lists its terms
// expand to (terms whose live df is 0 are skipped, query.rs:48):
//   0 lists  -> empty result          1 list -> class S: one DIRECT segment (seg_s[q])
//   >= 2     -> class G: qt_gcount[t] segments per query term, filled by gfill_kernel.
__global__ void plan_query_kernel(IndexView ix, uint64_t n_queries,
                                  const uint64_t* __restrict__ query_term_off,
                                  const uint32_t* __restrict__ qt_lo, const uint32_t* __restrict__ qt_hi,
                                  const uint32_t* __restrict__ qt_len, Seg* __restrict__ seg_s,
                                  unsigned long long* __restrict__ s_tiles, unsigned long long* __restrict__ qt_gcount,
                                  uint32_t* __restrict__ qt_q, unsigned long long* __restrict__ q_isg,
                                  unsigned long long* __restrict__ q_grows, unsigned long long* __restrict__ stats,
                                  UPlan up, uint64_t n_qterms) {
  uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (q == 0) {                // the exclusive scans run over n + 1 entries: give the extra input entry a defined value
    s_tiles[n_queries] = 0ull; qt_gcount[n_qterms] = 0ull;
    if (up.q_isu) { up.q_isu[n_queries] = 0ull; up.q_isu2[n_queries] = 0ull; }
  }
  unsigned long long st_rows = 0, st_live = 0, st_ptr = 0, st_compact = 0;
  unsigned long long su_rows = 0, su_live = 0, su_ptr = 0;
  if (q < n_queries) {
  uint64_t t0 = query_term_off[q], t1 = query_term_off[q + 1];
  uint32_t nl = 0;
  uint64_t rows = 0;
  uint64_t single = t0;
  for (uint64_t t = t0; t < t1; ++t) {
    uint32_t lo = qt_lo[t], hi = qt_hi[t];
    uint32_t c = ix.live_prefix[hi] - ix.live_prefix[lo];
    if (c) single = t;
    nl += c;
    rows += ix.liverows_prefix[hi] - ix.liverows_prefix[lo];
    qt_q[t] = (uint32_t)q;
  }
  Seg s;
  s.row_begin = 0; s.n_rows = 0; s.q = (uint32_t)q; s.term = 0; s.qlen = 0; s.qti = 0;
  s.mode = MODE_DIRECT; s.pad = 0; s.slot = 0;
  unsigned long long tiles = 0;
  if (nl == 1) {
    uint32_t term = first_live_term(ix, qt_lo[single], qt_hi[single]);
    uint64_t a = ix.term_row_begin[term], b = ix.term_row_begin[term + 1];
    s.row_begin = a; s.n_rows = (uint32_t)(b - a); s.term = term; s.qlen = qt_len[single];
    s.qti = (uint16_t)(single - t0);
    tiles = ((b + TILE_ROWS - 1) / TILE_ROWS) - (a / TILE_ROWS);
    // the launch's row statistics are known without touching a row:
    st_rows = b - a;                       // rows streamed
    st_live = ix.term_live_rows[term];     // rows whose doc is live = score() evaluations
    st_ptr = ix.term_df_live[term];        // reference DocumentPointer visits (sum of multiplicities)
    if (ix.cpost != nullptr && ix.term_compact[term]) {    // rows of the interior tiles stream from the compact copy
      const uint64_t i0 = (a + TILE_ROWS - 1) / TILE_ROWS, i1 = b / TILE_ROWS;
      if (i1 > i0) st_compact = (i1 - i0) * TILE_ROWS;
    }
  }
  seg_s[q] = s;
  s_tiles[q] = tiles;
  bool g = nl >= 2;
  // class U (union_kernels.cuh): ZeroToOne, several lists, a large part of the corpus, few enough
  // query terms / overlaps for the dense per-doc state
  bool u = false, ugen = false;
  if (up.enabled && g && rows >= up.min_rows && t1 - t0 <= 255) {
    UQuery d;
    d.q = (uint32_t)q; d.qtl = (uint32_t)(t1 - t0); d.n_act = 0; d.overlaps = 0;
    bool ok = true;
    for (uint64_t t = t0; t < t1 && ok; ++t) {
      const uint32_t lo = qt_lo[t], hi = qt_hi[t];
      if (ix.live_prefix[hi] == ix.live_prefix[lo]) continue;
      if (d.n_act == (uint32_t)U_MAX_ACT || hi - lo > (1u << 20) || qt_len[t] > 63u) { ok = false; break; }
      // the kernel's (explen - qlen, tf) table must cover every row the query can meet
      if (up.max_term_bytes - qt_len[t] >= (uint32_t)U_DE || up.max_tf >= (uint32_t)U_TF) { ok = false; break; }
      d.lo[d.n_act] = lo; d.hi[d.n_act] = hi; d.qti[d.n_act] = (uint8_t)(t - t0); d.qlen[d.n_act] = (uint8_t)qt_len[t];
      ++d.n_act;
    }
    for (uint32_t a = d.n_act; a < (uint32_t)U_MAX_ACT; ++a) { d.lo[a] = 0; d.hi[a] = 0; d.qti[a] = 0; d.qlen[a] = 0; }
    for (uint32_t a = 0; a < (uint32_t)U_MAX_ACT; ++a) {
      uint32_t dep = a < d.n_act ? 1u : 0u;          // + one candidate per other query term whose range overlaps
      for (uint32_t b2 = 0; b2 < d.n_act && a < d.n_act; ++b2)
        if (b2 != a && d.lo[a] < d.hi[b2] && d.lo[b2] < d.hi[a]) ++dep;
      d.depth[a] = (uint8_t)dep; d.pad[a] = 0;
      if (dep > 1) d.overlaps = 1;
      ok = ok && dep <= (up.allow_gen ? 2u : 1u);    // the general kernel keeps two candidates per query term
    }
    if (ok) {
      u = true;
      ugen = d.overlaps != 0;
      g = false;
      up.uq[q] = d;
      for (uint32_t a = 0; a < d.n_act; ++a) {
        su_rows += ix.term_row_begin[d.hi[a]] - ix.term_row_begin[d.lo[a]];      // the union image streams every row of the range
        su_live += up.liverowcnt_prefix[d.hi[a]] - up.liverowcnt_prefix[d.lo[a]];
        su_ptr += up.dflive_prefix[d.hi[a]] - up.dflive_prefix[d.lo[a]];
      }
    }
  }
  if (up.q_isu) { up.q_isu[q] = (u && !ugen) ? 1ull : 0ull; up.q_isu2[q] = (u && ugen) ? 1ull : 0ull; }
  q_isg[q] = g ? 1ull : 0ull;
  q_grows[q] = g ? rows : 0ull;
  for (uint64_t t = t0; t < t1; ++t)
    qt_gcount[t] = g ? (unsigned long long)(ix.live_prefix[qt_hi[t]] - ix.live_prefix[qt_lo[t]]) : 0ull;
  }
  st_rows = warp_sum_u64(st_rows); st_live = warp_sum_u64(st_live); st_ptr = warp_sum_u64(st_ptr);
  st_compact = warp_sum_u64(st_compact);
  if ((threadIdx.x & 31) == 0 && st_rows) {
    if (st_compact) atomicAdd(&stats[ST_ROWS_COMPACT], st_compact);
    atomicAdd(&stats[ST_ROWS_STREAMED], st_rows);
    atomicAdd(&stats[ST_ROWS_SCORED], st_live);
    atomicAdd(&stats[ST_POINTER_VISITS], st_ptr);
  }
  su_rows = warp_sum_u64(su_rows); su_live = warp_sum_u64(su_live); su_ptr = warp_sum_u64(su_ptr);
  if ((threadIdx.x & 31) == 0 && su_rows) {
    atomicAdd(&up.stats[ST_ROWS_STREAMED], su_rows);
    atomicAdd(&up.stats[ST_ROWS_SCORED], su_live);
    atomicAdd(&up.stats[ST_POINTER_VISITS], su_ptr);
  }
}

// class-U query indices, compacted in query order (q_uidx = exclusive prefix of q_isu)
__global__ void ucompact_kernel(uint64_t n_queries, const unsigned long long* __restrict__ q_isu,
                                const unsigned long long* __restrict__ q_uidx, uint32_t* __restrict__ u_list) {
  uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (q < n_queries && q_isu[q]) u_list[q_uidx[q]] = (uint32_t)q;
}

// ---- union image build (pb_index_create): rows re-sorted by (doc shard, term, doc) --------------------
// pass 1: one warp per term writes, for each of its rows, the sort key (shard) and remembers the term
__global__ void ubuild_keys_kernel(IndexView ix, uint32_t wbits, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                   uint32_t* __restrict__ row_term) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t t = warp; t < ix.n_terms; t += nwarps) {
    const uint64_t a = ix.term_row_begin[t], b = ix.term_row_begin[t + 1];
    for (uint64_t r = a + lane; r < b; r += 32) {
      keys[r] = row_doc(ix, r) >> wbits;
      vals[r] = (uint32_t)r;
      row_term[r] = (uint32_t)t;
    }
  }
}
// pass 2 (after a stable radix sort of (shard, row)): gather the rows into the union columns
__global__ void ubuild_gather_kernel(IndexView ix, uint32_t wbits, uint64_t n_rows, const uint32_t* __restrict__ sorted_row,
                                     const uint32_t* __restrict__ row_term, uint32_t* __restrict__ u_meta,
                                     uint32_t* __restrict__ u_term, uint16_t* __restrict__ c0, uint16_t* __restrict__ c1,
                                     uint16_t* __restrict__ c2, uint16_t* __restrict__ c3) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  const uint64_t r = sorted_row[i];
  const uint32_t t = row_term[r], doc = row_doc(ix, r);
  u_meta[i] = (doc & ((1u << wbits) - 1u)) | (ix.term_byte_len[t] << 16);
  u_term[i] = t;
  uint16_t* cs[4] = {c0, c1, c2, c3};
  for (uint32_t f = 0; f < ix.num_fields; ++f)
    cs[f][i] = (uint16_t)((row_col(ix, r, (int)f) << ix.fl_bits[f]) | row_col(ix, r, (int)(ix.num_fields + f)));
}
// first row of every shard in the sorted key array (keys ascend)
__global__ void ubuild_bounds_kernel(const uint32_t* __restrict__ sorted_keys, uint64_t n_rows, uint32_t n_shards,
                                     uint32_t* __restrict__ shard_row) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > n_shards) return;
  uint64_t lo = 0, hi = n_rows;
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if (sorted_keys[mid] < s) lo = mid + 1; else hi = mid;
  }
  shard_row[s] = (uint32_t)lo;
}

// ---- device-side flatten of the posting columns (pb_index_create_from_builder) -----------------
// The host builder keeps an append log of (term id, doc, tf[F]) tuples in document order.  Pass 1: sort key = the
// term's DFS ordinal (+ the exact maxima of tf / field length that decide the column layout).  A stable radix sort
// of (key, tuple index) then puts the rows of a term together with their docs ascending; pass 2 writes the tile-blocked
// columns (u16 codes or u32 columns) straight into HBM: the host never materialises them.
struct FlatTuple { uint32_t term, doc, tf[4]; };
__global__ void flat_keys_kernel(const FlatTuple* __restrict__ t, uint64_t n, const uint32_t* __restrict__ ord_of,
                                 const uint32_t* __restrict__ doc_fl, uint32_t F, uint32_t* __restrict__ keys,
                                 uint32_t* __restrict__ vals, uint32_t* __restrict__ maxima) {
  uint32_t mtf[4] = {0, 0, 0, 0}, mfl[4] = {0, 0, 0, 0};
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const FlatTuple tp = t[i];
    keys[i] = ord_of[tp.term];
    vals[i] = (uint32_t)i;
    for (uint32_t f = 0; f < F; ++f) { mtf[f] = max(mtf[f], tp.tf[f]); mfl[f] = max(mfl[f], doc_fl[(uint64_t)tp.doc * F + f]); }
  }
  for (uint32_t f = 0; f < F; ++f) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mtf[f] = max(mtf[f], __shfl_xor_sync(0xffffffffu, mtf[f], o));
      mfl[f] = max(mfl[f], __shfl_xor_sync(0xffffffffu, mfl[f], o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMax(&maxima[f], mtf[f]); atomicMax(&maxima[4 + f], mfl[f]); }
  }
}
__global__ void flat_gather_kernel(const FlatTuple* __restrict__ t, const uint32_t* __restrict__ sorted_idx, uint64_t n,
                                   const uint32_t* __restrict__ doc_fl, uint32_t F, uint32_t narrow, uint32_t tile_words,
                                   uint32_t fb0, uint32_t fb1, uint32_t fb2, uint32_t fb3, uint32_t* __restrict__ post) {
  const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  const FlatTuple tp = t[sorted_idx[r]];
  const uint32_t fl_bits[4] = {fb0, fb1, fb2, fb3};
  uint32_t* tile = post + (r / TILE_ROWS) * (uint64_t)tile_words;
  const uint32_t in = (uint32_t)(r % TILE_ROWS);
  tile[in] = tp.doc;
  if (narrow) {
    uint16_t* codes = reinterpret_cast<uint16_t*>(tile + TILE_ROWS);
    for (uint32_t f = 0; f < F; ++f)
      codes[f * TILE_ROWS + in] = (uint16_t)((tp.tf[f] << fl_bits[f]) | doc_fl[(uint64_t)tp.doc * F + f]);
  } else {
    for (uint32_t f = 0; f < F; ++f) {
      tile[(1 + f) * TILE_ROWS + in] = tp.tf[f];
      tile[(1 + F + f) * TILE_ROWS + in] = doc_fl[(uint64_t)tp.doc * F + f];
    }
  }
}

// One warp per query term of a class-G query: writes one SECONDARY segment per live expanded
// term, in expansion order, and elects the query's largest list (atomicMax on rows<<32|seg).
__global__ void gfill_kernel(IndexView ix, uint64_t n_qterms, const uint64_t* __restrict__ query_term_off,
                             const uint32_t* __restrict__ qt_lo, const uint32_t* __restrict__ qt_hi,
                             const uint32_t* __restrict__ qt_len, const uint32_t* __restrict__ qt_q,
                             const unsigned long long* __restrict__ qt_gcount, const unsigned long long* __restrict__ qt_goff,
                             Seg* __restrict__ seg_g, unsigned long long* __restrict__ g_tiles,
                             unsigned long long* __restrict__ q_prim, unsigned long long* __restrict__ stats) {
  int lane = threadIdx.x & 31;
  uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  unsigned long long st_rows = 0, st_live = 0, st_ptr = 0;
  for (uint64_t t = warp; t < n_qterms; t += nwarps) {
    if (qt_gcount[t] == 0) continue;
    uint32_t lo = qt_lo[t], hi = qt_hi[t], q = qt_q[t];
    uint32_t qlen = qt_len[t];
    uint16_t qti = (uint16_t)(t - query_term_off[q]);
    uint64_t out = qt_goff[t];
    unsigned long long best = 0;
    for (uint32_t base = lo; base < hi; base += 32) {
      uint32_t term = base + lane;
      bool live = term < hi && ix.term_df_live[term] > 0;
      uint32_t m = __ballot_sync(0xffffffffu, live);
      if (live) {
        uint64_t idx = out + __popc(m & ((1u << lane) - 1u));
        uint64_t a = ix.term_row_begin[term], b = ix.term_row_begin[term + 1];
        Seg s;
        s.row_begin = a; s.n_rows = (uint32_t)(b - a); s.q = q; s.term = term; s.qlen = qlen;
        s.qti = qti; s.mode = MODE_SECONDARY; s.pad = 0; s.slot = 0;
        seg_g[idx] = s;
        g_tiles[idx] = ((b + TILE_ROWS - 1) / TILE_ROWS) - (a / TILE_ROWS);
        unsigned long long cand = ((unsigned long long)(b - a) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)idx);
        best = max(best, cand);
        st_rows += b - a; st_live += ix.term_live_rows[term]; st_ptr += ix.term_df_live[term];
      }
      out += __popc(m);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = max(best, (unsigned long long)shfl_u64(best, lane ^ o));
    if (lane == 0 && best) atomicMax(&q_prim[q], best);
  }
  st_rows = warp_sum_u64(st_rows); st_live = warp_sum_u64(st_live); st_ptr = warp_sum_u64(st_ptr);
  if (lane == 0 && st_rows) {
    atomicAdd(&stats[ST_ROWS_STREAMED], st_rows);
    atomicAdd(&stats[ST_ROWS_SCORED], st_live);
    atomicAdd(&stats[ST_POINTER_VISITS], st_ptr);
  }
}

// One thread per query: promote the elected list to PRIMARY and bound the side-path records:
// every secondary row + at most one primary row per secondary doc.
__global__ void gprimary_kernel(uint64_t n_queries, uint32_t doc_bits, const unsigned long long* __restrict__ q_isg,
                                const unsigned long long* __restrict__ q_grows,
                                const unsigned long long* __restrict__ q_prim, Seg* __restrict__ seg_g,
                                unsigned long long* __restrict__ q_recbound, unsigned long long* __restrict__ q_nbins,
                                uint8_t* __restrict__ q_scheme, uint8_t* __restrict__ q_shift,
                                const uint64_t* __restrict__ query_term_off,
                                const unsigned long long* __restrict__ qt_goff,
                                unsigned long long* __restrict__ q_gsegoff, unsigned long long* __restrict__ q_bmwords,
                                uint32_t bitmap_words, unsigned long long* __restrict__ qmax) {
  uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (q > n_queries) return;
  q_gsegoff[q] = qt_goff[query_term_off[q]];   // qt_goff has n_qterms + 1 entries
  if (q == n_queries) return;
  unsigned long long bound = 0, nbins = 0, bmw = 0;
  uint8_t scheme = 0, shift = 0;
  if (q_isg[q]) {
    unsigned long long p = q_prim[q];
    uint32_t idx = 0xFFFFFFFFu - (uint32_t)(p & 0xFFFFFFFFull);
    unsigned long long prows = p >> 32;
    unsigned long long rows = q_grows[q], secondary = rows - prows;
    if (secondary * 16ull >= rows) {
      // exact scheme: the record count is only known after the marking pass; plan with an estimate
      scheme = 1;
      bound = rows / 2 + 1024;
      bmw = bitmap_words;
    } else {
      // primary scheme: every secondary row + at most one primary row per secondary doc
      seg_g[idx].mode = MODE_PRIMARY;
      bound = 2ull * secondary;
      // row mask: 4 words per tile the primary list touches
      const unsigned long long a = seg_g[idx].row_begin, e = a + seg_g[idx].n_rows;
      bmw = 4ull * ((e + TILE_ROWS - 1) / TILE_ROWS - a / TILE_ROWS) + 12ull;     // + pads: the scoring loop reads masks 2 tiles ahead
    }
    // doc-range bins of width 2^shift sized for ~8 records each (a warp window holds 32)
    const unsigned long long n_docs_pow = 1ull << doc_bits;
    unsigned long long want = bound / 8 + 1;                 // number of bins wanted
    uint32_t sh = doc_bits;
    while (sh > 0 && (n_docs_pow >> sh) < want) --sh;
    shift = (uint8_t)sh;
    nbins = (n_docs_pow >> sh);
  }
  q_recbound[q] = bound;
  q_bmwords[q] = bmw;
  if (bound) { atomicMax(&qmax[0], bound); atomicMax(&qmax[1], bmw); }     // largest single query (host: capacities)
  q_nbins[q] = nbins;
  q_scheme[q] = scheme;
  q_shift[q] = shift;
}

// Per round: slot = rank of the query among the round's class-G queries (key of the sorted fallback).
// Also writes the tile count of every list the MARKING pass walks (all but the primary lists): its
// prefix sum is the marking pass's own tile space, so that its warps share that work evenly.
__global__ void gslot_kernel(Seg* __restrict__ seg_g, uint64_t seg_begin, uint64_t seg_end,
                             const unsigned long long* __restrict__ q_gidx, const uint8_t* __restrict__ q_scheme,
                             uint32_t q_begin, const unsigned long long* __restrict__ g_tiles,
                             unsigned long long* __restrict__ g_mtiles) {
  uint64_t i = seg_begin + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i > seg_end) return;
  if (i == seg_end) { g_mtiles[i] = 0ull; return; }
  const uint32_t q = seg_g[i].q;
  seg_g[i].slot = (uint32_t)(q_gidx[q] - q_gidx[q_begin]);
  if (q_scheme[q]) seg_g[i].mode = MODE_MULTI;
  g_mtiles[i] = seg_g[i].mode == MODE_PRIMARY ? 0ull : g_tiles[i];
}

// One warp per query: merge the partial top-k lists into the final (score desc, doc asc) top-k.
__global__ void __launch_bounds__(CTA_THREADS) finalize_kernel(Outputs o, uint64_t n_queries) {
  const int lane = threadIdx.x & 31;
  const uint64_t w = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t W = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  if (o.k == 0) return;
  for (uint64_t q = w; q < n_queries; q += W) {
    uint32_t p = o.part_head[q];
    if (p == NONE) continue;
    WarpAcc acc;
    acc.reset((uint32_t)q);
    while (p != NONE) {
      uint32_t n = o.part_n[p];
      bool v = lane < (int)n;
      uint32_t d = v ? o.part_doc[(size_t)p * o.k + lane] : NONE;
      double s = v ? o.part_score[(size_t)p * o.k + lane] : -1.0;
      acc.insert_candidates(v && better(s, d, acc.thr_s, acc.thr_d), d, s, lane, (int)o.k);
      p = o.part_next[p];
    }
    uint32_t ntop = (uint32_t)min((unsigned long long)o.k, o.n_results[q]);
    if (lane < (int)ntop) {
      o.topk_doc[(size_t)q * o.k + lane] = acc.td;
      o.topk_score[(size_t)q * o.k + lane] = acc.ts;
    }
    if (lane == 0) o.topk_n[q] = ntop;
  }
}

// ------------------------------------------------------------------------------------------
// Compact copy of the narrow tiles (IndexView::cpost): one warp per FULL tile.  A tile whose docs span less than
// 2^16 gets its doc column as u16 offsets from the tile's smallest doc, its code columns copied behind, and its
// base in cbase[tile]; any other tile gets cbase = NONE and its block is left zero.
// ------------------------------------------------------------------------------------------
__global__ void compact_tiles_kernel(const uint32_t* __restrict__ post, uint32_t tile_words, uint32_t F, uint64_t n_full_tiles,
                                     uint64_t n_tiles_alloc, uint32_t* __restrict__ cpost, uint32_t* __restrict__ cbase) {
  const int lane = threadIdx.x & 31;
  const uint64_t w = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t W = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint32_t CW = (TILE_ROWS / 2) * (1 + F);
  for (uint64_t t = w; t < n_tiles_alloc; t += W) {
    if (t >= n_full_tiles) { if (lane == 0) cbase[t] = NONE; continue; }
    const uint32_t* src = post + t * (uint64_t)tile_words;
    const uint4 d = *reinterpret_cast<const uint4*>(src + lane * 4);
    uint32_t lo = min(min(d.x, d.y), min(d.z, d.w)), hi = max(max(d.x, d.y), max(d.z, d.w));
#pragma unroll
    for (int of = 16; of > 0; of >>= 1) {
      lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, of));
      hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, of));
    }
    const bool fit = hi - lo < 65536u;
    if (lane == 0) cbase[t] = fit ? lo : NONE;
    if (!fit) continue;
    uint32_t* dst = cpost + t * (uint64_t)CW;
    *reinterpret_cast<uint2*>(dst + lane * 2) = make_uint2((d.x - lo) | ((d.y - lo) << 16), (d.z - lo) | ((d.w - lo) << 16));
    for (uint32_t f = 0; f < F; ++f)
      *reinterpret_cast<uint2*>(dst + (TILE_ROWS / 2) * (1 + f) + lane * 2) =
          *reinterpret_cast<const uint2*>(src + TILE_ROWS + f * (TILE_ROWS / 2) + lane * 2);
  }
}

// IndexView::row_dead: bit r % 128 of tile r / 128 = the doc of posting row r is removed.  One warp per tile (the pad rows of
// the last tile read doc 0: whatever they get is never used, the scoring loop masks rows outside a list).
__global__ void row_dead_kernel(const uint32_t* __restrict__ post, uint32_t tile_words, uint64_t n_tiles,
                                const uint32_t* __restrict__ removed, uint32_t n_docs, uint32_t* __restrict__ row_dead) {
  const int lane = threadIdx.x & 31;
  const uint64_t w = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t W = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t t = w; t < n_tiles; t += W) {
    const uint4 d = *reinterpret_cast<const uint4*>(post + t * (uint64_t)tile_words + lane * 4);
    const uint32_t dv[4] = {d.x, d.y, d.z, d.w};
    uint32_t nib = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (dv[j] < n_docs && ((__ldg(&removed[dv[j] >> 5]) >> (dv[j] & 31)) & 1u)) nib |= 1u << j;
    uint32_t word = nib << ((lane & 7) * 4);
    word |= __shfl_xor_sync(0xffffffffu, word, 1);
    word |= __shfl_xor_sync(0xffffffffu, word, 2);
    word |= __shfl_xor_sync(0xffffffffu, word, 4);
    if ((lane & 7) == 0) row_dead[t * 4 + (lane >> 3)] = word;
  }
}

// term_compact[t] = the list has at least `min_tiles` interior tiles and every one of them is compact.  One warp per term.
__global__ void term_compact_kernel(const uint64_t* __restrict__ term_row_begin, uint32_t n_terms, const uint32_t* __restrict__ cbase,
                                    uint32_t min_tiles, uint8_t* __restrict__ term_compact, unsigned long long* __restrict__ rows_compact) {
  const int lane = threadIdx.x & 31;
  const uint64_t w = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t W = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t t = w; t < n_terms; t += W) {
    const uint64_t a = term_row_begin[t], b = term_row_begin[t + 1];
    const uint64_t i0 = (a + TILE_ROWS - 1) / TILE_ROWS, i1 = b / TILE_ROWS;
    bool ok = i1 >= i0 + min_tiles;
    if (ok) {
      bool bad = false;
      for (uint64_t i = i0 + lane; i < i1; i += 32) bad = bad || cbase[i] == NONE;
      ok = !__any_sync(0xffffffffu, bad);
    }
    if (lane == 0) {
      term_compact[t] = ok ? 1 : 0;
      if (ok) atomicAdd(rows_compact, (unsigned long long)((i1 - i0) * TILE_ROWS));
    }
  }
}

}  // namespace pbk
