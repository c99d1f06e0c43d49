// The dense union path: ZeroToOne (src/score/default/zero_to_one.rs:44-126) for UNION-HEAVY queries —
// several query terms, each a prefix that expands to hundreds of posting lists, together touching a
// large part of the corpus (BASELINE cfg 2).  The per-list machinery of score_kernel / mark / binfold
// touches such rows three times; here every row is read ONCE.
//
// Layout ("union image", built on the device at pb_index_create next to the term-major image — HBM is
// plentiful, 8 + 2F bytes per row): the SAME posting rows re-sorted by (doc shard, term, doc), shard =
// 2^wbits consecutive doc ordinals.  Terms are numbered in DFS order, so ALL expansions of a query term
// (term range [lo, hi), query.rs:109-147) are ONE contiguous run of rows inside a shard — hundreds of
// short posting lists collapse into one coalesced stream per (query term, shard).
//   u_meta[r] = doc ordinal inside the shard | byte length of the expanded term << 16
//   u_term[r] = term ordinal;   u_code[f][r] = tf << fl_bits | fl  (the narrow image's code)
//
// Algorithm (one CTA per (doc chunk, query) item, dynamic item counter, chunk-major so that all CTAs sweep
// the same ~5 MB slice of the image at a time), per shard, three passes over the runs / the shard's docs:
//   A  count: every live row adds 1 to its doc's event counter in shared memory (reads u_meta only);
//   B  score: the rows stream by again (L1 / L2 hits).  A row whose doc has exactly ONE event IS the doc's
//      whole result: its ZeroToOne score is computed inline, four rows per lane, and consumed (count,
//      digests, top-k) exactly like the single-list kernel does — about 85 % of the result docs of cfg 2.
//      Rows of docs with several events feed a DENSE per-doc state in shared memory instead,
//        state[field][doc][query term] = min over the events of (expanded byte length, expansion rank, tf)
//      (one u32, shared-memory atomicMin);
//   C  merge: the docs with several events are collected (warp compaction) and one lane per doc replays
//      ZeroToOne::finalize on the candidates.
// Why a minimum is enough
// (zero_to_one.rs:98-121): finalize() sorts a (doc, field)'s entries by score descending (stable) and
// accepts, per query term, the FIRST entry whose term pool is not exhausted.  The entry score
// 1 - |explen - qlen| / explen falls strictly with explen for one query term, ties keep expansion order =
// term order, so "first" = minimum of (explen, rank).  An entry can only be refused by the pool when
// ANOTHER query term accepted the same expanded term before — possible only between query terms whose
// term ranges overlap (one is a prefix of the other).  Queries without such overlaps (97 % of cfg 2) keep
// one candidate per query term (GEN = false); with overlaps a query term keeps its best TWO candidates
// (GEN = true: it overlaps with at most one other query term, so one refusal is the most it can see).
// After the runs of a shard, one thread per doc replays finalize() on the candidates exactly (sort
// order, consumed query terms, pools, f64 operation order) and feeds count / digests / top-k like every
// other path.  Both divisions of zero_to_one.rs:117-120 are division-free here: (s / tf) * tf comes from a
// per-query table built with real divisions, the division by max(field_length, query_terms_len) is
// RN(1/m) from a table + one FMA correction step, which scripts/prove_z2o_rcp.c compares with the real
// quotient bit for bit over the WHOLE domain this path admits (term bytes <= 63, tf <= 63, m <= 255).
#pragma once
#include "kernels.cuh"

namespace pbk {

constexpr int U_MAX_ACT = 4;         // query terms with at least one live posting list
constexpr int U_DE = 10, U_TF = 16;  // table of (s / tf) * tf by (explen - qlen, tf); outside -> computed
#ifndef PB_U_CHUNK
#define PB_U_CHUNK 32
#endif
constexpr int U_CHUNK = PB_U_CHUNK;  // shards per work item
constexpr uint32_t U_SENT = 0xFFFFFFFFu;
#ifndef PB_U_THREADS
#define PB_U_THREADS 256
#endif
#ifndef PB_U_MINBLOCKS
#define PB_U_MINBLOCKS 2
#endif
template <bool GEN> struct UShape { static constexpr int THREADS = GEN ? 512 : PB_U_THREADS; };

struct __align__(16) UQuery {
  uint32_t q, qtl;                 // query index; query_terms_len (query.rs:32)
  uint32_t n_act, overlaps;        // live query terms; does any pair of their term ranges overlap
  uint32_t lo[U_MAX_ACT], hi[U_MAX_ACT];
  uint8_t qti[U_MAX_ACT], qlen[U_MAX_ACT], depth[U_MAX_ACT], pad[U_MAX_ACT];
};
static_assert(sizeof(UQuery) == 64, "UQuery layout");

struct UnionView {
  const uint32_t* meta;
  const uint32_t* term;
  const uint16_t* code[4];
  const uint32_t* shard_row;       // [n_shards + 1]
  uint32_t n_shards, wbits;
  uint32_t ok;                     // the image exists (eligibility: see pb_index_create)
};

struct UnionParams {
  IndexView ix;
  UnionView uv;
  Outputs out;
  const UQuery* uq;                // [n_queries], valid where the query is class U
  const uint32_t* u_list;          // class-U query indices of this launch
  uint32_t n_u, n_chunks;
  unsigned long long* item_counter;
  unsigned long long* prof;        // -DPB_UNION_PROF=1: SM cycles by phase {setup, stream, finalize, merge, items}
};
#ifndef PB_UNION_PROF
#define PB_UNION_PROF 0
#endif

__host__ __device__ inline size_t union_smem_bytes(int F, uint32_t wbits, bool gen) {
  const size_t W = (size_t)1 << wbits;
  return (size_t)F * W * 16 * (gen ? 2 : 1) + (size_t)F * W + (size_t)W * 8    // candidate records, field lengths, event counters x2
         + (size_t)U_MAX_ACT * U_DE * U_TF * 8 + (size_t)U_MAX_ACT * U_DE * 8  // v table, s table
         + 256 * 16                                                            // {m, RN(1/m)}
         + (size_t)U_CHUNK * U_MAX_ACT * 2 * 4 + (size_t)W * 4 + 16;           // run bounds, multi-event doc lists x2
}

// The rows of a shard are read twice (count pass, score pass) a few microseconds apart: these loads ALLOCATE in L1
// (the single-list kernel streams with L1::no_allocate), and the lines of the next pass / next shard are prefetched.
__device__ __forceinline__ uint4 ldg_keep(const uint32_t* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ldg_keep_u64(const uint32_t* p) {
  uint2 r;
  asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// one 128-row group of the union image = 4 lines of meta, 4 of term, 2 per code column: lane l prefetches line l
template <int F>
__device__ __forceinline__ void prefetch_group(const UnionView& uv, uint32_t grp, int lane, bool with_meta) {
  const uint32_t base = grp << 7;
  if (with_meta && lane < 4) prefetch_l1(uv.meta + base + lane * 32);
  else if (lane >= 4 && lane < 8) prefetch_l1(uv.term + base + (lane - 4) * 32);
  else if (lane >= 8 && lane < 8 + 2 * F) prefetch_l1(uv.code[(lane - 8) >> 1] + base + ((lane - 8) & 1) * 64);
}

// max of two doubles that are >= +0.0 and finite: such doubles order like their bit patterns, so this is fmax without
// the NaN / signed-zero handling an f64 max instruction sequence carries
__device__ __forceinline__ double u_max_nonneg(double a, double b) {
  const long long ia = __double_as_longlong(a), ib = __double_as_longlong(b);
  return __longlong_as_double(ia > ib ? ia : ib);
}

__device__ __forceinline__ uint32_t u4get(const uint4& v, int a) { return a == 0 ? v.x : a == 1 ? v.y : a == 2 ? v.z : v.w; }

// (s / tf) * tf for a candidate key of query term `a` (zero_to_one.rs:72, 117-118)
__device__ __forceinline__ double u_entry_num(const double* __restrict__ vt, const UQuery& uq, int a, uint32_t key) {
  const uint32_t e = key >> 26, tf = key & 63u, ql = uq.qlen[a];
  const uint32_t de = e - ql;
  if (de < (uint32_t)U_DE && tf < (uint32_t)U_TF) return vt[(a * U_DE + de) * U_TF + tf];
  const double s = z2o_term_score(e, ql);
  return __dmul_rn(fmin(__ddiv_rn(s, (double)tf), 1.0), (double)tf);
}
// x / m through RN(1/m) and one correction step: bit-identical to __ddiv_rn on this path's domain
__device__ __forceinline__ double u_div_m(double x, double md, double y) {
  const double q = __dmul_rn(x, y);
  const double r = __fma_rn(-q, md, x);
  return __fma_rn(r, y, q);
}

// The exact replay of finalize() for one (doc, field) with up to 8 candidates (tops then seconds): used when
// three or more entries are summed (the order of the f64 additions matters) or a term pool may refuse an entry.
static __device__ __noinline__ double u_finalize_slow(const double* __restrict__ vt, const double* __restrict__ stab,
                                                      const UQuery& uq, uint4 top, uint4 sec, double md, double y) {
  uint32_t ck[8];
  ck[0] = top.x; ck[1] = sec.x; ck[2] = top.y; ck[3] = sec.y; ck[4] = top.z; ck[5] = sec.z; ck[6] = top.w; ck[7] = sec.w;
  // candidates enumerated in (query term, expansion rank) order = the insertion order of the reference
  double sv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sv[i] = -1.0;
    if (ck[i] != U_SENT) {
      const uint32_t e = ck[i] >> 26, ql = uq.qlen[i >> 1];
      sv[i] = (e - ql) < (uint32_t)U_DE ? stab[(i >> 1) * U_DE + (e - ql)] : z2o_term_score(e, ql);   // zero_to_one.rs:72
    }
  }
  double accx = 0.0;
  uint32_t done = 0, consumed = 0, accepted = 0;
#pragma unroll 1
  for (int step = 0; step < 8; ++step) {
    int best = -1;
    double bs = -1.0;
    uint32_t bkey = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (!((done >> i) & 1u) && sv[i] > bs) { best = i; bs = sv[i]; bkey = ck[i]; }   // strict >: ties stay in insertion order
    if (best < 0) break;
    done |= 1u << best;
    const int ba = best >> 1;
    if ((consumed >> ba) & 1u) continue;                                  // zero_to_one.rs:101
    const uint32_t tf = bkey & 63u;
    const uint32_t bterm = uq.lo[ba] + ((bkey >> 6) & 0xFFFFFu);
    uint32_t used = 0;                                                    // zero_to_one.rs:104-113: the term's pool
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if ((accepted >> i) & 1u) used += (uq.lo[i >> 1] + ((ck[i] >> 6) & 0xFFFFFu)) == bterm ? 1u : 0u;
    if (used >= tf) continue;
    consumed |= 1u << ba;
    accepted |= 1u << best;
    accx = __dadd_rn(accx, u_div_m(u_entry_num(vt, uq, ba, bkey), md, y));   // zero_to_one.rs:117-120
  }
  return accx;
}

template <int F, bool GEN>
__global__ void __launch_bounds__(UShape<GEN>::THREADS, GEN ? 1 : (F <= 2 ? PB_U_MINBLOCKS : 1))
union_kernel(const __grid_constant__ UnionParams P) {
  constexpr int THREADS = UShape<GEN>::THREADS;
  constexpr int NW = THREADS / 32;
  extern __shared__ __align__(16) unsigned char u_smem[];
  const uint32_t W = 1u << P.uv.wbits;
  uint4* top = reinterpret_cast<uint4*>(u_smem);                                  // [F][W]: best key per query term
  uint4* sec = top + (GEN ? (size_t)F * W : 0);                                   // [F][W]: second best (GEN)
  uint32_t* cnt = reinterpret_cast<uint32_t*>(top + (size_t)F * W * (GEN ? 2 : 1)); // [2][W] events per doc (two shards in flight)
  double2* mrc = reinterpret_cast<double2*>(cnt + 2 * W);                         // [256] {m, RN(1 / m)}
  double* vt = reinterpret_cast<double*>(mrc + 256);                              // [act][U_DE][U_TF] (s / tf) * tf
  double* stab = vt + U_MAX_ACT * U_DE * U_TF;                                    // [act][U_DE] s
  uint32_t* bnd = reinterpret_cast<uint32_t*>(stab + U_MAX_ACT * U_DE);           // [U_CHUNK][act][2]
  uint8_t* flv = reinterpret_cast<uint8_t*>(bnd + U_CHUNK * U_MAX_ACT * 2);       // [F][W]
  __shared__ UQuery uq;
  __shared__ unsigned long long s_item;
  __shared__ unsigned long long red_dd[NW], red_sd[NW];
  __shared__ uint32_t red_cnt[NW];
  __shared__ uint32_t n_multi[2];           // docs of the current shard with several events (two counters, alternating) ...
  uint16_t* mlist = reinterpret_cast<uint16_t*>(flv + (size_t)F * W);   // ... and their lists [2][W]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned long long n_items = (unsigned long long)P.n_u * P.n_chunks;
  const uint4 SENT4 = make_uint4(U_SENT, U_SENT, U_SENT, U_SENT);
  const bool capture = P.out.full_q != nullptr;

  for (uint32_t i = tid; i < 256; i += THREADS) mrc[i] = make_double2((double)i, __ldg(&P.ix.rcp[i]));
  for (uint32_t i = tid; i < 2 * W; i += THREADS) cnt[i] = 0u;
  if (tid == 0) { n_multi[0] = 0u; n_multi[1] = 0u; }
#if PB_UNION_PROF
  long long pc[6] = {0, 0, 0, 0, 0, 0}, pt = clock64();
#define PB_UPROF(i) do { if (tid == 0) { const long long n_ = clock64(); pc[i] += n_ - pt; pt = n_; } } while (0)
#else
#define PB_UPROF(i) do {} while (0)
#endif
  for (;;) {
    __syncthreads();                                   // everything of the previous item is consumed
    PB_UPROF(3);
    if (tid == 0) s_item = atomicAdd(P.item_counter, 1ull);
    __syncthreads();
    const unsigned long long item = s_item;
    if (item >= n_items) break;
    const uint32_t chunk = (uint32_t)(item / P.n_u);
    const uint32_t qsel = P.u_list[(uint32_t)(item % P.n_u)];
    if (tid < (int)(sizeof(UQuery) / 4)) reinterpret_cast<uint32_t*>(&uq)[tid] = reinterpret_cast<const uint32_t*>(&P.uq[qsel])[tid];
    __syncthreads();
    const uint32_t n_act = uq.n_act, qtl = uq.qtl;
    const uint32_t s0 = chunk * U_CHUNK, s1 = min(P.uv.n_shards, s0 + U_CHUNK);
    // ---- per item: run bounds of every (shard, query term), the (s, v) tables, a clean state ----------
    for (uint32_t i = tid; i < (s1 - s0) * n_act * 2; i += THREADS) {
      const uint32_t sh = i / (n_act * 2), a = (i >> 1) % n_act;
      const uint32_t target = (i & 1) ? uq.hi[a] : uq.lo[a];
      uint32_t lo = P.uv.shard_row[s0 + sh], hi = P.uv.shard_row[s0 + sh + 1];
      while (lo < hi) {                                 // first row of the shard whose term is >= target
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(&P.uv.term[mid]) < target) lo = mid + 1; else hi = mid;
      }
      bnd[(sh * U_MAX_ACT + a) * 2 + (i & 1)] = lo;
    }
    for (uint32_t i = tid; i < n_act * U_DE * U_TF; i += THREADS) {
      const uint32_t a = i / (U_DE * U_TF), de = (i / U_TF) % U_DE, tf = i % U_TF;
      const uint32_t ql = uq.qlen[a];
      const double s = z2o_term_score(ql + de, ql);                       // zero_to_one.rs:72
      if (tf == 0) stab[a * U_DE + de] = s;
      vt[i] = tf ? __dmul_rn(fmin(__ddiv_rn(s, (double)tf), 1.0), (double)tf) : 0.0;   // zero_to_one.rs:117-118
    }
    for (uint32_t i = tid; i < (uint32_t)F * W * (GEN ? 2u : 1u); i += THREADS) top[i] = SENT4;
    if (tid == 0) { n_multi[0] = 0u; n_multi[1] = 0u; }
    WarpAcc acc;
    acc.reset(uq.q);
    __syncthreads();
    PB_UPROF(0);
#if PB_UNION_PROF
    if (tid == 0) ++pc[4];
#endif

    // The shards of the item run as a two-barrier software pipeline:
    //   [score(s)]  barrier  [merge(s) and count(next shard) side by side]  barrier  [score(next)] ...
    // count(next) only touches the OTHER event-counter / list buffer, so it shares a barrier interval with merge(s).
    struct Geo { uint32_t o1, o2, o3, total, ga, gb, gc, gd; };
    auto geo_of = [&](uint32_t sh) {
      const uint32_t* b = bnd + (sh - s0) * U_MAX_ACT * 2;
      uint32_t n[U_MAX_ACT], g[U_MAX_ACT];
#pragma unroll
      for (int a = 0; a < U_MAX_ACT; ++a) {
        n[a] = 0; g[a] = 0;
        if (a < (int)n_act && b[a * 2 + 1] > b[a * 2]) { g[a] = b[a * 2] >> 7; n[a] = ((b[a * 2 + 1] + 127u) >> 7) - g[a]; }
      }
      Geo G;
      G.o1 = n[0]; G.o2 = G.o1 + n[1]; G.o3 = G.o2 + n[2]; G.total = G.o3 + n[3];
      G.ga = g[0]; G.gb = g[1]; G.gc = g[2]; G.gd = g[3];
      return G;
    };
    auto next_shard = [&](uint32_t from, Geo& G) {        // first shard >= from in which the query has rows (CTA-uniform)
      for (uint32_t sh = from; sh < s1; ++sh) {
        G = geo_of(sh);
        if (G.total) return sh;
      }
      return s1;
    };

    // ---- pass A: events per doc ---------------------------------------------------------------------
    auto pass_count = [&](uint32_t sh, const Geo& G, uint32_t buf) {
      const uint32_t* b = bnd + (sh - s0) * U_MAX_ACT * 2;
      uint32_t* cn = cnt + buf * W;
      uint16_t* ml = mlist + buf * W;
      const uint32_t doc_base = sh << P.uv.wbits;
      for (uint32_t t = warp; t < G.total; t += NW) {
        const int a = (t >= G.o1 ? 1 : 0) + (t >= G.o2 ? 1 : 0) + (t >= G.o3 ? 1 : 0);
        const uint32_t grp = a == 0 ? G.ga + t : a == 1 ? G.gb + (t - G.o1) : a == 2 ? G.gc + (t - G.o2) : G.gd + (t - G.o3);
        const uint32_t r0 = b[a * 2], r1 = b[a * 2 + 1];
        const uint32_t base = (grp << 7) + lane * 4;
        const uint4 m4 = ldg_keep(P.uv.meta + base);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t r = base + j;
          if (r < r0 || r >= r1) continue;
          const uint32_t dl = u4c(m4, j) & 0xFFFFu;
          if (P.ix.has_removed) {                       // removed-but-not-vacuumed docs are skipped (query.rs:65)
            const uint32_t d = doc_base + dl;
            if ((__ldg(&P.ix.removed[d >> 5]) >> (d & 31)) & 1u) continue;
          }
          if (atomicAdd(&cn[dl], 1u) == 1u) ml[atomicAdd(&n_multi[buf], 1u)] = (uint16_t)dl;   // second event: the doc needs the merge
        }
      }
    };

    // ---- pass B: score the single-event docs inline, feed the dense state with the rest -------------
    auto pass_score = [&](uint32_t sh, const Geo& G, uint32_t buf) {
      const uint32_t* b = bnd + (sh - s0) * U_MAX_ACT * 2;
      uint32_t* cn = cnt + buf * W;
      const uint32_t doc_base = sh << P.uv.wbits;
      for (uint32_t t = warp; t < G.total; t += NW) {
        const int a = (t >= G.o1 ? 1 : 0) + (t >= G.o2 ? 1 : 0) + (t >= G.o3 ? 1 : 0);
        const uint32_t grp = a == 0 ? G.ga + t : a == 1 ? G.gb + (t - G.o1) : a == 2 ? G.gc + (t - G.o2) : G.gd + (t - G.o3);
        const uint32_t r0 = b[a * 2], r1 = b[a * 2 + 1];
        const uint32_t base = (grp << 7) + lane * 4;
        const uint4 m4 = ldg_keep(P.uv.meta + base);
        const uint4 t4 = ldg_keep(P.uv.term + base);
        uint2 c4[F];
#pragma unroll
        for (int f = 0; f < F; ++f) c4[f] = ldg_keep_u64(reinterpret_cast<const uint32_t*>(P.uv.code[f] + base));
        const uint32_t lo_a = uq.lo[a], ql = uq.qlen[a];
        const bool two = GEN && uq.depth[a] > 1;
        const double* vta = vt + a * (U_DE * U_TF);
        uint32_t some = 0, multi = 0, dv[4];
        double sc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t r = base + j;
          const uint32_t meta = u4c(m4, j);
          const uint32_t dl = meta & 0xFFFFu, de = ((meta >> 16) & 0xFFu) - ql;
          const uint32_t c = (r >= r0 && r < r1) ? cn[dl] : 0u;     // 0: outside the run or a removed doc
          dv[j] = doc_base + dl;
          some |= (c == 1u) ? (1u << j) : 0u;
          multi |= (c > 1u) ? (1u << j) : 0u;
          if (c == 1u) cn[dl] = 0u;                     // the doc's only row: clean for the shard after next
          // zero_to_one.rs:44-126 for a doc with ONE event: max over the fields the term occurs in of
          // (s / tf) * tf / max(field_length, query_terms_len); table rows for tf = 0 hold +0.0
          double best = 0.0;
#pragma unroll
          for (int f = 0; f < F; ++f) {
            const uint32_t code = __byte_perm(j < 2 ? c4[f].x : c4[f].y, 0u, (j & 1) ? 0x4432u : 0x4410u);
            const uint32_t tf = code >> P.ix.fl_bits[f], fl = code & ((1u << P.ix.fl_bits[f]) - 1u);
            // class-U queries are only those whose (explen - qlen, tf) pairs all lie inside the table (plan_query_kernel);
            // rows outside the run (c == 0) may index anywhere inside the table block: de is clamped
            const double v = vta[min(de, (uint32_t)U_DE - 1u) * U_TF + tf];
            const double2 my = mrc[max(fl, qtl)];
            best = u_max_nonneg(u_div_m(v, my.x, my.y), best);
          }
          sc[j] = best;
        }
        if (multi) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (!((multi >> j) & 1u)) continue;
            const uint32_t meta = u4c(m4, j);
            const uint32_t dl = meta & 0xFFFFu;
            const uint32_t kbase = (((meta >> 16) & 0xFFu) << 26) | ((u4c(t4, j) - lo_a) << 6);
#pragma unroll
            for (int f = 0; f < F; ++f) {
              const uint32_t code = __byte_perm(j < 2 ? c4[f].x : c4[f].y, 0u, (j & 1) ? 0x4432u : 0x4410u);
              const uint32_t tf = code >> P.ix.fl_bits[f];
              if (tf == 0) continue;                    // zero_to_one.rs:56: only fields the term occurs in
              flv[f * W + dl] = (uint8_t)(code & ((1u << P.ix.fl_bits[f]) - 1u));
              uint32_t v = kbase | tf;
              const uint32_t old = atomicMin(reinterpret_cast<uint32_t*>(top + (f * W + dl)) + a, v);
              if (two) {                                // keep the two smallest keys: the loser moves on
                v = max(old, v);
                if (v != U_SENT) atomicMin(reinterpret_cast<uint32_t*>(sec + (f * W + dl)) + a, v);
              }
            }
          }
        }
        if (capture) acc.template add4<true, true>(P.out, some, dv, sc, lane);
        else acc.template add4<false, true>(P.out, some, dv, sc, lane);
      }
    };

    // ---- pass C: ZeroToOne::finalize (zero_to_one.rs:84-126) for the docs with several events: 32 listed docs
    //      at a time per warp, one lane per doc.
    auto pass_merge = [&](uint32_t sh, uint32_t buf) {
      uint32_t* cn = cnt + buf * W;
      const uint16_t* ml = mlist + buf * W;
      const uint32_t doc_base = sh << P.uv.wbits;
      const uint32_t nm = n_multi[buf];
      // the warps take batches from the top so that the count pass of the next shard (taken from warp 0 up) meets them
      for (uint32_t base = (NW - 1 - warp) * 32; base < nm; base += NW * 32) {
        const bool valid = base + lane < nm;
        const uint32_t d = valid ? ml[base + lane] : 0u;
        double result = 0.0;
        if (valid) {
          cn[d] = 0u;                                                      // clean for the shard after next
#pragma unroll
          for (int f = 0; f < F; ++f) {
            const uint4 k = top[f * W + d];
            if ((k.x & k.y & k.z & k.w) == U_SENT) continue;
            top[f * W + d] = SENT4;
            uint4 k2 = SENT4;
            if (GEN) { k2 = sec[f * W + d]; sec[f * W + d] = SENT4; }
            const double2 my = mrc[max((uint32_t)flv[f * W + d], qtl)];     // zero_to_one.rs:119
            uint32_t pm = (k.x != U_SENT ? 1u : 0u) | (k.y != U_SENT ? 2u : 0u) | (k.z != U_SENT ? 4u : 0u) | (k.w != U_SENT ? 8u : 0u);
            bool pool = false;
            if (GEN && (pm & (pm - 1u))) {
              // a pool can only refuse an entry when two query terms hold the SAME expanded term
              uint32_t tm[U_MAX_ACT];
#pragma unroll
              for (int a = 0; a < U_MAX_ACT; ++a) tm[a] = uq.lo[a] + ((u4get(k, a) >> 6) & 0xFFFFFu);
#pragma unroll
              for (int a = 0; a < U_MAX_ACT; ++a)
#pragma unroll
                for (int c = a + 1; c < U_MAX_ACT; ++c)
                  pool |= ((pm >> a) & (pm >> c) & 1u) && tm[a] == tm[c];
            }
            double accx = 0.0;
            if (GEN && pool) {
              accx = u_finalize_slow(vt, stab, uq, k, k2, my.x, my.y);
            } else if (__popc(pm) <= 2) {
              // <= 2 entries, no pool interaction: every query term accepts its best entry; a + b is commutative
              // and 0.0 + x == x, so no ordering is needed
              while (pm) {
                const int a = __ffs(pm) - 1;
                pm &= pm - 1u;
                accx = __dadd_rn(accx, u_div_m(u_entry_num(vt, uq, a, u4get(k, a)), my.x, my.y));
              }
            } else {
              // 3 or 4 entries: (x + y) + z — added in the order of the reference's stable sort by score descending;
              // absent terms sort last and add an exact +0.0
              double s4[U_MAX_ACT], en[U_MAX_ACT];
#pragma unroll
              for (int a = 0; a < U_MAX_ACT; ++a) {
                const uint32_t key = u4get(k, a);
                s4[a] = -1.0; en[a] = 0.0;
                if (key != U_SENT) {
                  const uint32_t e = key >> 26, ql = uq.qlen[a];
                  s4[a] = (e - ql) < (uint32_t)U_DE ? stab[a * U_DE + (e - ql)] : z2o_term_score(e, ql);
                  en[a] = u_div_m(u_entry_num(vt, uq, a, key), my.x, my.y);
                }
              }
              int rk[U_MAX_ACT];
#pragma unroll
              for (int a = 0; a < U_MAX_ACT; ++a) {
                rk[a] = 0;
#pragma unroll
                for (int c = 0; c < U_MAX_ACT; ++c)
                  if (c != a) rk[a] += (s4[c] > s4[a] || (s4[c] == s4[a] && c < a)) ? 1 : 0;
              }
#pragma unroll
              for (int r = 0; r < U_MAX_ACT; ++r)
                accx = __dadd_rn(accx, rk[0] == r ? en[0] : rk[1] == r ? en[1] : rk[2] == r ? en[2] : en[3]);
            }
            result = fmax(accx, result);                                    // zero_to_one.rs:122
          }
        }
        acc.add(P.out, valid, doc_base + d, result, lane);
      }
    };

    {
      Geo G, Gn;
      uint32_t buf = 0;
      uint32_t cur = next_shard(s0, G);
      if (cur < s1) { pass_count(cur, G, buf); __syncthreads(); }
      PB_UPROF(1);
      while (cur < s1) {
        if (tid == 0) n_multi[buf ^ 1u] = 0u;              // last read two barriers ago (merge of the previous shard)
        pass_score(cur, G, buf);
        __syncthreads();
        PB_UPROF(2);
        const uint32_t nxt = next_shard(cur + 1, Gn);
        pass_merge(cur, buf);
        if (nxt < s1) pass_count(nxt, Gn, buf ^ 1u);
        __syncthreads();
        PB_UPROF(5);
        cur = nxt; G = Gn; buf ^= 1u;
      }
    }

    // ---- item end: one partial result (count, digests, top-k) per item --------------------------------
    {
      uint32_t total = acc.cnt;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
      const uint64_t tdd = warp_sum_u64(acc.dd), tsd = warp_sum_u64(acc.sd);
      double* m_ts = reinterpret_cast<double*>(u_smem);                  // the state area is free now
      uint32_t* m_td = reinterpret_cast<uint32_t*>(m_ts + NW * 32);
      if (lane == 0) { red_cnt[warp] = total; red_dd[warp] = tdd; red_sd[warp] = tsd; }
      m_ts[warp * 32 + lane] = acc.ts;
      m_td[warp * 32 + lane] = acc.td;
      __syncthreads();
      if (warp == 0) {
        WarpAcc m;
        m.reset(uq.q);
        uint32_t c = lane < NW ? red_cnt[lane] : 0u;
        uint64_t dd = lane < NW ? red_dd[lane] : 0ull, sd = lane < NW ? red_sd[lane] : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        dd = warp_sum_u64(dd); sd = warp_sum_u64(sd);
        if (c) {
          if (P.out.k) {
            for (int w2 = 0; w2 < NW; ++w2) {
              const double cs = m_ts[w2 * 32 + lane];
              const uint32_t cd = m_td[w2 * 32 + lane];
              m.insert_candidates(cd != NONE && better(cs, cd, m.thr_s, m.thr_d), cd, cs, lane, (int)P.out.k);
            }
          }
          m.cnt = lane == 0 ? c : 0u;
          m.dd = lane == 0 ? dd : 0ull;
          m.sd = lane == 0 ? sd : 0ull;
          m.flush(P.out, false, lane);
        }
      }
    }
  }
#if PB_UNION_PROF
  if (tid == 0 && P.prof)
    for (int i = 0; i < 6; ++i) atomicAdd(&P.prof[i], (unsigned long long)pc[i]);
#endif
}

}  // namespace pbk
