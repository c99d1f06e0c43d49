// union_warp_kernel — the dense union path with WARP-AUTONOMOUS tasks (see union_kernels.cuh for the algorithm and
// for why a per-(doc, field, query term) minimum reproduces ZeroToOne::finalize, zero_to_one.rs:84-126).
//
// union_kernel runs a 2048-doc shard per CTA: three short passes separated by CTA barriers, 10-14 tiles over 8 warps —
// ncu shows 36 % issue utilisation with 3.4 of 4 warps per scheduler parked at a barrier.  Here the shard is small
// (2^wbits = 512 docs) and ONE WARP owns a (query, shard) task from its first row to its last result: the three passes
// are separated by __syncwarp only, tasks are handed out inside the CTA by a shared-memory counter, and 24 warps per SM
// each work on an independent task — nothing waits for a sibling.
//   - rows are processed one per lane, the runs of the query's terms concatenated, so the lanes stay full whatever
//     the run lengths are;
//   - the dense state is kept only for the docs with several events: a doc's second event gives it a slot
//     (slot[doc]); a task with more than U_W_MAXM such docs replays the score pass once per block of slots (the rows
//     come from L1 by then), so shared memory per warp stays at a few KB whatever the query;
//   - the CTA still shares the per-query tables and the run bounds of the item's shards, and merges its warps'
//     accumulators once per item.
#pragma once
#include "union_kernels.cuh"

namespace pbk {

constexpr int U_W_THREADS = 256;
constexpr int U_W_NW = U_W_THREADS / 32;
constexpr int U_W_MAXM = 96;             // multi-event docs resolved per replay of the score pass
constexpr uint32_t U_ITEM_DOCS = 65536;  // docs per (chunk, query) work item

__host__ __device__ inline size_t union_warp_smem_bytes(int F, uint32_t wbits, bool gen) {
  const size_t W = (size_t)1 << wbits, spi = U_ITEM_DOCS >> wbits;
  const size_t per_warp = W * 2 /*cnt*/ + W * 2 /*slot*/ + W * 2 /*mlist*/ + (size_t)F * U_W_MAXM * 16 * (gen ? 2 : 1) + (size_t)F * U_W_MAXM;
  return 256 * 16 + (size_t)U_MAX_ACT * U_DE * U_TF * 8 + (size_t)U_MAX_ACT * U_DE * 8 + spi * U_MAX_ACT * 2 * 4 +
         ((per_warp + 15) & ~(size_t)15) * U_W_NW + 32;
}

template <int F, bool GEN>
__global__ void __launch_bounds__(U_W_THREADS, GEN ? 2 : 3) union_warp_kernel(const __grid_constant__ UnionParams P) {
  extern __shared__ __align__(16) unsigned char w_smem[];
  const uint32_t W = 1u << P.uv.wbits, spi = U_ITEM_DOCS >> P.uv.wbits;
  double2* mrc = reinterpret_cast<double2*>(w_smem);                              // [256] {m, RN(1 / m)}
  double* vt = reinterpret_cast<double*>(mrc + 256);                              // [act][U_DE][U_TF] (s / tf) * tf
  double* stab = vt + U_MAX_ACT * U_DE * U_TF;                                    // [act][U_DE] s
  uint32_t* bnd = reinterpret_cast<uint32_t*>(stab + U_MAX_ACT * U_DE);           // [spi][act][2]
  const size_t per_warp = ((size_t)W * 6 + (size_t)F * U_W_MAXM * 16 * (GEN ? 2 : 1) + (size_t)F * U_W_MAXM + 15) & ~(size_t)15;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned char* mine = reinterpret_cast<unsigned char*>(bnd + spi * U_MAX_ACT * 2) + per_warp * warp;
  uint4* top = reinterpret_cast<uint4*>(mine);                                    // [F][MAXM] best key per query term
  uint4* sec = top + (GEN ? F * U_W_MAXM : 0);                                    // [F][MAXM] second best (GEN)
  uint16_t* cnt = reinterpret_cast<uint16_t*>(top + F * U_W_MAXM * (GEN ? 2 : 1)); // [W] events per doc
  uint16_t* slot = cnt + W;                                                       // [W] slot of a multi-event doc
  uint16_t* mlist = slot + W;                                                     // [W] the multi-event docs
  uint8_t* flv = reinterpret_cast<uint8_t*>(mlist + W);                           // [F][MAXM]
  __shared__ UQuery uq;
  __shared__ unsigned long long s_item;
  __shared__ uint32_t s_next;                 // next shard of the item (handed out to the warps)
  __shared__ uint32_t s_nm[U_W_NW];           // per warp: multi-event docs of its current task
  __shared__ unsigned long long red_dd[U_W_NW], red_sd[U_W_NW];
  __shared__ uint32_t red_cnt[U_W_NW];
  __shared__ double m_ts[U_W_NW * 32];
  __shared__ uint32_t m_td[U_W_NW * 32];

  const unsigned long long n_items = (unsigned long long)P.n_u * P.n_chunks;
  const uint4 SENT4 = make_uint4(U_SENT, U_SENT, U_SENT, U_SENT);
  for (uint32_t i = tid; i < 256; i += U_W_THREADS) mrc[i] = make_double2((double)i, __ldg(&P.ix.rcp[i]));
  for (uint32_t i = lane; i < W; i += 32) cnt[i] = 0;
  if (lane == 0) s_nm[warp] = 0u;

  for (;;) {
    __syncthreads();                                   // everything of the previous item is consumed
    if (tid == 0) { s_item = atomicAdd(P.item_counter, 1ull); s_next = 0u; }
    __syncthreads();
    const unsigned long long item = s_item;
    if (item >= n_items) break;
    const uint32_t chunk = (uint32_t)(item / P.n_u);
    const uint32_t qsel = P.u_list[(uint32_t)(item % P.n_u)];
    if (tid < (int)(sizeof(UQuery) / 4)) reinterpret_cast<uint32_t*>(&uq)[tid] = reinterpret_cast<const uint32_t*>(&P.uq[qsel])[tid];
    __syncthreads();
    const uint32_t n_act = uq.n_act, qtl = uq.qtl;
    const uint32_t s0 = chunk * spi, s1 = min(P.uv.n_shards, s0 + spi);
    // ---- per item: run bounds of every (shard, query term) and the (s, v) tables ------------------------
    for (uint32_t i = tid; i < (s1 - s0) * n_act * 2; i += U_W_THREADS) {
      const uint32_t sh = i / (n_act * 2), a = (i >> 1) % n_act;
      const uint32_t target = (i & 1) ? uq.hi[a] : uq.lo[a];
      uint32_t lo = P.uv.shard_row[s0 + sh], hi = P.uv.shard_row[s0 + sh + 1];
      while (lo < hi) {                                 // first row of the shard whose term is >= target
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(&P.uv.term[mid]) < target) lo = mid + 1; else hi = mid;
      }
      bnd[(sh * U_MAX_ACT + a) * 2 + (i & 1)] = lo;
    }
    for (uint32_t i = tid; i < n_act * U_DE * U_TF; i += U_W_THREADS) {
      const uint32_t a = i / (U_DE * U_TF), de = (i / U_TF) % U_DE, tf = i % U_TF;
      const uint32_t ql = uq.qlen[a];
      const double s = z2o_term_score(ql + de, ql);                       // zero_to_one.rs:72
      if (tf == 0) stab[a * U_DE + de] = s;
      vt[i] = tf ? __dmul_rn(fmin(__ddiv_rn(s, (double)tf), 1.0), (double)tf) : 0.0;   // zero_to_one.rs:117-118
    }
    WarpAcc acc;
    acc.reset(uq.q);
    __syncthreads();

    // ---- the warp's tasks: one (query, shard) each, start to finish ---------------------------------------
    for (;;) {
      uint32_t shl = 0;
      if (lane == 0) shl = atomicAdd(&s_next, 1u);
      shl = __shfl_sync(0xffffffffu, shl, 0);
      if (shl >= s1 - s0) break;
      const uint32_t* b = bnd + shl * U_MAX_ACT * 2;
      // the runs of the query's terms, concatenated into one row space
      uint32_t r0a = 0, r0b = 0, r0c = 0, r0d = 0, o1, o2, o3, total;
      {
        uint32_t n[U_MAX_ACT], r0[U_MAX_ACT];
#pragma unroll
        for (int a = 0; a < U_MAX_ACT; ++a) {
          n[a] = 0; r0[a] = 0;
          if (a < (int)n_act && b[a * 2 + 1] > b[a * 2]) { r0[a] = b[a * 2]; n[a] = b[a * 2 + 1] - b[a * 2]; }
        }
        o1 = n[0]; o2 = o1 + n[1]; o3 = o2 + n[2]; total = o3 + n[3];
        r0a = r0[0]; r0b = r0[1]; r0c = r0[2]; r0d = r0[3];
      }
      if (total == 0) continue;
      const uint32_t doc_base = (s0 + shl) << P.uv.wbits;
      auto row_of = [&](uint32_t i, int& a) {
        a = (i >= o1 ? 1 : 0) + (i >= o2 ? 1 : 0) + (i >= o3 ? 1 : 0);
        return a == 0 ? r0a + i : a == 1 ? r0b + (i - o1) : a == 2 ? r0c + (i - o2) : r0d + (i - o3);
      };

      // ---- pass A: events per doc; a doc's second event gives it a slot ----------------------------------
      for (uint32_t i0 = lane; i0 < total; i0 += 128) {
        // four steps' loads in flight before the first one is used; the score pass reads term / code columns of the
        // same rows a few hundred cycles later: their lines are pulled towards L1 meanwhile
        uint32_t mv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t i = i0 + 32 * u;
          mv[u] = 0;
          if (i < total) {
            int a;
            const uint32_t r = row_of(i, a);
            mv[u] = __ldg(&P.uv.meta[r]);
            if ((lane & 7) == 0) {
              prefetch_l1(P.uv.term + r);
#pragma unroll
              for (int f = 0; f < F; ++f) prefetch_l1(P.uv.code[f] + r);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (i0 + 32 * u >= total) continue;
          const uint32_t dl = mv[u] & 0xFFFFu;
          if (P.ix.has_removed) {                         // removed-but-not-vacuumed docs are skipped (query.rs:65)
            const uint32_t d = doc_base + dl;
            if ((__ldg(&P.ix.removed[d >> 5]) >> (d & 31)) & 1u) continue;
          }
          // two u16 counters share a word: a doc can receive at most 4 * 255 * F events (field lengths <= 255), far below 2^16
          const uint32_t sh16 = (dl & 1u) * 16u;
          const uint32_t old = (atomicAdd(reinterpret_cast<uint32_t*>(cnt) + (dl >> 1), 1u << sh16) >> sh16) & 0xFFFFu;
          if (old == 1u) {
            const uint32_t pos = atomicAdd(&s_nm[warp], 1u);
            mlist[pos] = (uint16_t)dl;
            slot[dl] = (uint16_t)pos;
          }
        }
      }
      __syncwarp();
      const uint32_t nm = s_nm[warp];

      // ---- pass B (+ C) once per block of U_W_MAXM multi-event docs; the first replay also scores the single-event docs
      for (uint32_t slot0 = 0; slot0 == 0 || slot0 < nm; slot0 += U_W_MAXM) {
        const uint32_t n_here = min((uint32_t)U_W_MAXM, nm > slot0 ? nm - slot0 : 0u);
        for (uint32_t i = lane; i < (uint32_t)F * n_here * (GEN ? 2u : 1u); i += 32) {
          // top[f][0 .. n_here), then sec[f][0 .. n_here)
          const uint32_t half = i / ((uint32_t)F * n_here), j = i % ((uint32_t)F * n_here);
          (half ? sec : top)[(j / n_here) * U_W_MAXM + (j % n_here)] = SENT4;
        }
        __syncwarp();
        // the loads of a row (doc, term, codes) do not depend on each other: issued together, one step ahead
        struct RowRegs { uint32_t meta, term, code[F]; int a; };
        auto load_row = [&](uint32_t i, RowRegs& R) {
          R.meta = 0; R.term = 0; R.a = 0;
#pragma unroll
          for (int f = 0; f < F; ++f) R.code[f] = 0;
          if (i < total) {
            const uint32_t r = row_of(i, R.a);
            R.meta = __ldg(&P.uv.meta[r]);
            R.term = __ldg(&P.uv.term[r]);
#pragma unroll
            for (int f = 0; f < F; ++f) R.code[f] = __ldg(&P.uv.code[f][r]);
          }
        };
        RowRegs nxt;
        load_row(lane, nxt);
        for (uint32_t i0 = 0; i0 < total; i0 += 32) {
          const uint32_t i = i0 + lane;
          const RowRegs cur = nxt;
          load_row(i + 32, nxt);
          bool single = false;
          uint32_t doc = 0;
          double best = 0.0;
          if (i < total) {
            const int a = cur.a;
            const uint32_t meta = cur.meta;
            const uint32_t dl = meta & 0xFFFFu, e = (meta >> 16) & 0xFFu;
            const uint32_t c = cnt[dl];                  // 0: a removed doc
            doc = doc_base + dl;
            uint32_t code[F];
#pragma unroll
            for (int f = 0; f < F; ++f) code[f] = cur.code[f];
            if (c == 1u) {
              if (slot0 == 0) {
                // zero_to_one.rs:44-126 for a doc with ONE event: max over the fields the term occurs in of
                // (s / tf) * tf / max(field_length, query_terms_len); table rows for tf = 0 hold +0.0
                single = true;
                cnt[dl] = 0;                             // the doc's only row: clean for the next task
                const double* vta = vt + a * (U_DE * U_TF) + min(e - uq.qlen[a], (uint32_t)U_DE - 1u) * U_TF;
#pragma unroll
                for (int f = 0; f < F; ++f) {
                  const uint32_t tf = code[f] >> P.ix.fl_bits[f], fl = code[f] & ((1u << P.ix.fl_bits[f]) - 1u);
                  const double2 my = mrc[max(fl, qtl)];
                  best = u_max_nonneg(u_div_m(vta[tf], my.x, my.y), best);
                }
              }
            } else if (c > 1u) {
              const uint32_t sl = (uint32_t)slot[dl] - slot0;
              if (sl < (uint32_t)U_W_MAXM) {
                const uint32_t kbase = (e << 26) | ((cur.term - uq.lo[a]) << 6);
                const bool two = GEN && uq.depth[a] > 1;
#pragma unroll
                for (int f = 0; f < F; ++f) {
                  const uint32_t tf = code[f] >> P.ix.fl_bits[f];
                  if (tf == 0) continue;                 // zero_to_one.rs:56: only fields the term occurs in
                  flv[f * U_W_MAXM + sl] = (uint8_t)(code[f] & ((1u << P.ix.fl_bits[f]) - 1u));
                  uint32_t v = kbase | tf;
                  const uint32_t old = atomicMin(reinterpret_cast<uint32_t*>(top + f * U_W_MAXM + sl) + a, v);
                  if (two) {                             // keep the two smallest keys: the loser moves on
                    v = max(old, v);
                    if (v != U_SENT) atomicMin(reinterpret_cast<uint32_t*>(sec + f * U_W_MAXM + sl) + a, v);
                  }
                }
              }
            }
          }
          if (slot0 == 0) acc.add_nonneg(P.out, single, doc, best, lane);
        }
        __syncwarp();
        // ---- pass C: ZeroToOne::finalize for this block's multi-event docs, one lane per doc -------------
        for (uint32_t j0 = 0; j0 < n_here; j0 += 32) {
          const uint32_t j = j0 + lane;
          const bool valid = j < n_here;
          const uint32_t d = valid ? mlist[slot0 + j] : 0u;
          double result = 0.0;
          if (valid) {
            cnt[d] = 0;                                  // clean for the next task
#pragma unroll
            for (int f = 0; f < F; ++f) {
              const uint4 k = top[f * U_W_MAXM + j];
              if ((k.x & k.y & k.z & k.w) == U_SENT) continue;
              uint4 k2 = SENT4;
              if (GEN) k2 = sec[f * U_W_MAXM + j];
              const double2 my = mrc[max((uint32_t)flv[f * U_W_MAXM + j], qtl)];   // zero_to_one.rs:119
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
                // <= 2 entries, no pool interaction: a + b is commutative and 0.0 + x == x, so no ordering is needed
                while (pm) {
                  const int a = __ffs(pm) - 1;
                  pm &= pm - 1u;
                  accx = __dadd_rn(accx, u_div_m(u_entry_num(vt, uq, a, u4get(k, a)), my.x, my.y));
                }
              } else {
                // 3 or 4 entries: added in the order of the reference's stable sort by score descending
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
        __syncwarp();
      }
      if (lane == 0) s_nm[warp] = 0u;
      __syncwarp();
    }

    // ---- item end: one partial result (count, digests, top-k) per item --------------------------------
    {
      uint32_t total = acc.cnt;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
      const uint64_t tdd = warp_sum_u64(acc.dd), tsd = warp_sum_u64(acc.sd);
      if (lane == 0) { red_cnt[warp] = total; red_dd[warp] = tdd; red_sd[warp] = tsd; }
      m_ts[warp * 32 + lane] = acc.ts;
      m_td[warp * 32 + lane] = acc.td;
      __syncthreads();
      if (warp == 0) {
        WarpAcc m;
        m.reset(uq.q);
        uint32_t c = lane < U_W_NW ? red_cnt[lane] : 0u;
        uint64_t dd = lane < U_W_NW ? red_dd[lane] : 0ull, sd = lane < U_W_NW ? red_sd[lane] : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        dd = warp_sum_u64(dd); sd = warp_sum_u64(sd);
        if (c) {
          if (P.out.k) {
            for (int w2 = 0; w2 < U_W_NW; ++w2) {
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
}

}  // namespace pbk
