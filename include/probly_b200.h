/*
 * probly_b200.h — C ABI of the B200-native query hot path of probly-search.
 *
 * This is the drop-in seam UNDER the reference's `Index::query` (src/query.rs:21-106): the
 * reference has no FFI of its own (crate-type cdylib but no #[no_mangle] item anywhere,
 * Cargo.toml:25-26), so these entry points are what a Rust shim (rust/src/lib.rs in this repo,
 * unbuilt — no rustc in the image) binds with `extern "C"`.  Plain pointers and sizes only; no
 * torch / C++ types cross the boundary.  See INTEGRATION.md for the binding stubs.
 *
 * Ownership: the caller owns every buffer it passes in or receives results in; the library
 * copies what it needs and owns only its opaque handles (+ their host/device memory).
 * Errors: 0 = ok, negative = error code; nothing throws or aborts across the ABI; a
 * thread-local message is available from pb_last_error().  (The reference panics instead:
 * `unwrap()` at src/query.rs:46,63,70 and the NaN sort at :103.)
 * Threading: a pb_builder needs external exclusion for mutation (mirrors `&mut self`); a
 * pb_index is immutable between pb_index_create / pb_index_set_live_state / pb_index_set_df_extra
 * calls, which need the same exclusion against running batches; each pb_batch owns its workspace
 * and stream, so DIFFERENT pb_batch objects on one index may run concurrently from different host
 * threads (mirrors `query(&self, ..)`, src/query.rs:22).  The one-call forms pb_query_batch /
 * pb_query_full / pb_index_expand_term borrow one of a few internal batches per index (PB_QUERY_SLOTS, default 4:
 * own stream and workspace each), so calls from different host threads overlap on the device; they exclude the
 * calls that change the index (pb_index_set_live_state, pb_index_set_df_extra, pb_index_attach_delta), which wait
 * for running one-call queries and hold new ones back.  pb_index_last_stats = the call that finished last.
 * A staged pb_batch follows pb_index_set_live_state: its BM25 table is rebuilt when the index's
 * live state changed since it was staged.
 *
 * There is NO CPU fallback: every query entry point fails with PB_ERR_NO_DEVICE when no
 * CUDA device is usable.
 */
#ifndef PROBLY_B200_H
#define PROBLY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB_OK 0
#define PB_ERR_INVALID (-1)      /* bad argument / malformed input */
#define PB_ERR_CUDA (-2)         /* a CUDA runtime call failed */
#define PB_ERR_NO_DEVICE (-3)    /* no usable CUDA device (there is no CPU fallback) */
#define PB_ERR_CAPACITY (-4)     /* caller-provided output buffer too small */
#define PB_ERR_UNSUPPORTED (-5)  /* outside the supported envelope (see limits below) */
#define PB_ERR_DUPLICATE_KEY (-6)
#define PB_ERR_NOMEM (-7)

/* ScoreCalculator implementations that exist on the device (src/score/default/{bm25,zero_to_one}.rs).
 * Any other `impl ScoreCalculator` is arbitrary host code and cannot run here. */
#define PB_SCORER_BM25 0u
#define PB_SCORER_ZERO_TO_ONE 1u

#define PB_MAX_FIELDS 4u   /* Index::new(fields_num) with fields_num <= 4 */
#define PB_MAX_TOP_K 32u   /* per-query top-k kept by the batch entry point */
#define PB_MAX_QUERY_TERMS 4095u

typedef struct pb_builder pb_builder; /* host-side mutable index: the L1 of src/index.rs:19-33 */
typedef struct pb_index pb_index;     /* immutable flattened image resident in HBM on one device */
typedef struct pb_batch pb_batch;     /* one uploaded query batch + its device workspace */

/* ------------------------------------------------------------------------------------------
 * Host-side index maintenance (replaces Index::new / add_document / remove_document / vacuum,
 * src/index.rs:37-60, 77-158, 161-191, 194-241 — as semantics; the data layout is new).
 * Tokens arrive PRE-TOKENIZED: field accessors (src/lib.rs:11) and the tokenizer
 * (src/lib.rs:14) are user function pointers and stay in the binding layer.
 * ---------------------------------------------------------------------------------------- */

/* One document's tokens.  Tokens of all values of all fields are concatenated in
 * (field, value, token) order; empty tokens are passed through (they are skipped exactly
 * where src/index.rs:101 skips them). */
typedef struct pb_doc_tokens {
  const uint8_t* tok_bytes;          /* UTF-8 bytes of all tokens, concatenated */
  const uint64_t* tok_off;           /* [n_tokens + 1] byte offsets into tok_bytes */
  const uint32_t* value_tok_count;   /* one entry per field VALUE: its token count */
  const uint32_t* field_value_count; /* [num_fields]: how many values the accessor returned */
} pb_doc_tokens;

int pb_builder_create(uint32_t num_fields, pb_builder** out);          /* Index::new, index.rs:37 */
void pb_builder_destroy(pb_builder* b);
/* Deviation from the reference, stated on purpose: adding a key that is live, or removed but not yet vacuumed,
 * returns PB_ERR_INVALID (the reference accepts it and leaves two posting chains for one key - undefined results,
 * SURVEY rule 13).  After pb_builder_vacuum the key may be added again. */
int pb_builder_add_document(pb_builder* b, uint64_t key, const pb_doc_tokens* doc); /* index.rs:77 */
/* Bulk add: n_docs documents, every field has exactly one value; field_tok_count is
 * [n_docs * num_fields]; tokens concatenated in (doc, field, token) order. */
int pb_builder_add_documents(pb_builder* b, uint64_t n_docs, const uint64_t* keys,
                             const uint8_t* tok_bytes, const uint64_t* tok_off,
                             const uint32_t* field_tok_count);
int pb_builder_remove_document(pb_builder* b, uint64_t key);           /* index.rs:161 (lazy) */
int pb_builder_vacuum(pb_builder* b);                                  /* index.rs:194 */

typedef struct pb_builder_info {
  uint32_t num_fields;
  uint64_t n_live_docs;       /* docs.len() */
  uint64_t n_doc_ordinals;    /* ordinals ever assigned and still referenced */
  uint64_t n_removed_pending; /* removed but not yet vacuumed */
  uint64_t n_terms;           /* trie nodes that own at least one posting */
  uint64_t n_nodes;           /* live trie nodes, root included (count_nodes, index.rs:464-480) */
  uint64_t n_rows;            /* de-duplicated (term, doc) posting rows */
  uint64_t n_pointers;        /* reference DocumentPointer count = sum of multiplicities */
  uint64_t field_sum[PB_MAX_FIELDS]; /* FieldDetails.sum, index.rs:391-396 */
  double field_avg[PB_MAX_FIELDS];   /* FieldDetails.avg */
} pb_builder_info;
int pb_builder_get_info(const pb_builder* b, pb_builder_info* out);

/* The flattened, immutable image: what lives in HBM.  All pointers are HOST pointers owned
 * by the builder and stay valid until the builder is mutated or destroyed.
 *   - trie: nodes renumbered in DFS pre-order with children in the reference's linked-list
 *     order (most recently created child first, index.rs:409-419), so the expansions of a
 *     prefix (query.rs:109-147) are the contiguous, already ordered term range
 *     [node_term_lo, node_term_hi).  Edges of a node are stored sorted by char (lookup only).
 *   - postings: one row per (term, doc), rows of a term contiguous, terms in DFS order, docs
 *     ascending by ordinal inside a term.  Columns doc / tf[0..F) / fl[0..F) are stored
 *     TILE-BLOCKED: rows are cut into tiles of 128 and a tile is one contiguous block of
 *     (1 + 2F) x 128 u32 = [doc x128][tf0 x128]..[tfF-1 x128][fl0 x128]..[flF-1 x128], so a
 *     tile is a single 512(1+2F)-byte stream (one TMA bulk copy) and every column slice of it
 *     is a coalesced 512 B line group.  Row r, column c lives at
 *     post_blocks[((r / 128) * (1 + 2F) + c) * 128 + r % 128].  The row space is padded to
 *     whole tiles plus one spare tile (pad rows are zero).
 *   - live state: removed-but-not-vacuumed bitmap, live doc count, per-field average. */
typedef struct pb_index_image {
  uint32_t version;     /* = 1 */
  uint32_t num_fields;
  uint64_t n_nodes, n_edges, n_terms, n_rows, n_rows_padded, n_docs;
  uint32_t max_term_bytes;
  uint32_t max_tf[PB_MAX_FIELDS];
  uint32_t max_fl[PB_MAX_FIELDS];
  const uint32_t* node_edge_begin; /* [n_nodes + 1] */
  const uint32_t* node_term_lo;    /* [n_nodes] */
  const uint32_t* node_term_hi;    /* [n_nodes] */
  const uint32_t* node_parent;     /* [n_nodes] (host-only: to rebuild term strings) */
  const uint32_t* node_char;       /* [n_nodes] Unicode scalar on the edge into the node */
  const uint32_t* edge_char;       /* [n_edges] sorted ascending inside each node */
  const uint32_t* edge_child;      /* [n_edges] */
  const uint64_t* term_row_begin;  /* [n_terms + 1] */
  const uint32_t* term_byte_len;   /* [n_terms] UTF-8 byte length (bm25.rs:51-52, zero_to_one.rs:57-58) */
  const uint32_t* term_node;       /* [n_terms] */
  const uint32_t* post_blocks;     /* [n_rows_padded / 128][1 + 2F][128] tile-blocked columns */
  const uint64_t* doc_key;         /* [n_docs] ordinal -> caller's key */
  const uint32_t* removed_bitmap;  /* [(n_docs + 31) / 32 + 1] words (one spare zero word), bit set = not live */
  uint64_t n_removed;
  uint64_t n_live_docs;
  double field_avg[PB_MAX_FIELDS];
} pb_index_image;
int pb_builder_flatten(pb_builder* b, pb_index_image* out);

/* Incremental maintenance (the reference's add_document, src/index.rs:77-158, is cheap and
 * this keeps it cheap): a DELTA segment is the image of the docs whose ordinal is >= from_doc_ordinal only —
 * the CURRENT trie and term order (so expansion order is the global one, query.rs:109-147) with just those
 * docs' posting rows.  A serving process keeps the big image resident, uploads the small delta image as a
 * second pb_index, gives each of the two the other's per-term live counts (pb_index_set_df_extra: the
 * reference has ONE posting list per term, so BM25's document frequency is the sum) and merges the
 * per-query results, which are disjoint by document.  pb_builder_flatten_term_ids returns, for the image
 * flattened last, the builder's stable term id of every term ordinal: the key that matches terms across
 * segments.  The builder hands out ONE image at a time: a later flatten reuses its buffers. */
int pb_builder_flatten_from(pb_builder* b, uint64_t from_doc_ordinal, pb_index_image* out);
int pb_builder_flatten_term_ids(const pb_builder* b, uint32_t* out, uint64_t cap, uint64_t* n_terms);

/* On-disk / wire format of an image (the reference has no serialisation at all: no serde, the index
 * lives only in RAM — SURVEY §5, §8f-2).  One file = header + section table + 64-byte aligned
 * sections + FNV-1a checksum over the header scalars and every section (layout: csrc/image_io.cpp).
 * pb_image_load validates magic, scalar ranges, section sizes against the scalars, the checksum and the
 * structure (monotone offsets, children / terms / doc ordinals in range, every (tf, field length) within
 * max_tf / max_fl — the same pass pb_index_create runs on any image); the returned handle owns the buffer the
 * image's pointers point into, so a serving process needs no pb_builder:
 *   pb_image_load(path, &f); pb_index_create(pb_image_file_image(f), dev, &ix); pb_image_file_free(f); */
typedef struct pb_image_file pb_image_file;
int pb_image_save(const pb_index_image* image, const char* path);
int pb_image_load(const char* path, pb_image_file** out);
const pb_index_image* pb_image_file_image(const pb_image_file* f);
void pb_image_file_free(pb_image_file* f);

/* ------------------------------------------------------------------------------------------
 * Device image
 * ---------------------------------------------------------------------------------------- */
int pb_device_count(void);
int pb_index_create(const pb_index_image* image, int device, pb_index** out);
/* GPU index construction: the posting columns are flattened ON THE DEVICE.  The host builder keeps
 * what it always kept — the trie, the term dictionary and an append log of (term, doc, tf) tuples in document
 * order (tokenising and hashing, src/index.rs:77-158) — and flattens only the small structures; the device sorts the
 * tuples by term ordinal (cub radix sort, stable: docs stay ascending), derives max tf / field length, picks the
 * layout and writes the tile-blocked columns where the scoring kernels read them.  from_doc_ordinal > 0 builds a
 * delta segment (pb_builder_flatten_from).  pb_builder_flatten_structure returns that image WITHOUT posting columns
 * (post_blocks = NULL, max_tf = max_fl = 0): node / term arrays, doc keys and the live state, which is all a caller
 * needs next to the pb_index. */
int pb_index_create_from_builder(pb_builder* b, uint64_t from_doc_ordinal, int device, pb_index** out);
int pb_builder_flatten_structure(pb_builder* b, uint64_t from_doc_ordinal, pb_index_image* out);
/* New removed set / N / avg after remove_document WITHOUT re-flattening (pre-vacuum state,
 * SURVEY §3.4 rule 9).  removed_ords is the FULL set of removed ordinals.  Recomputes the
 * per-term live occurrence count (count_documents, index.rs:282-297) on the device and the
 * idf table (bm25.rs:41-56) on the host with libm log. */
int pb_index_set_live_state(pb_index* ix, const uint32_t* removed_ords, uint64_t n_removed,
                            uint64_t n_live_docs, const double* field_avg);
/* Segmented index: live occurrence counts of every term of THIS image in the other segments ([n_terms], by term
 * ordinal; n = 0 clears).  BM25's idf (bm25.rs:41-56) is recomputed from local + extra counts. */
int pb_index_set_df_extra(pb_index* ix, const uint64_t* df_extra, uint64_t n);
/* The whole arrangement in one call: `delta` (a pb_index created from a pb_builder_flatten_from image of the same
 * builder, same device) becomes the delta segment of `ix`, which takes ownership (delta = NULL detaches and destroys it).
 * The term id arrays are pb_builder_flatten_term_ids of the two flattens.  From then on
 *   pb_query_batch / pb_query_full on `ix` answer for BOTH segments (counts and digests add, top-k lists merge);
 *   pb_index_set_live_state on `ix` takes the removed ordinals of the WHOLE index and keeps both segments and the
 *     document frequencies they exchange up to date;
 *   pb_batch_create / pb_index_expand_term refuse (PB_ERR_UNSUPPORTED): a staged batch runs on one image. */
int pb_index_attach_delta(pb_index* ix, pb_index* delta, const uint32_t* main_term_ids, uint64_t n_main_terms,
                          const uint32_t* delta_term_ids, uint64_t n_delta_terms);
void pb_index_destroy(pb_index* ix);
/* expand_term (query.rs:109-126) through the device descent kernel: the expansions of `term`
 * joined by '\n' into out (cap bytes).  *n_expansions / *needed are always set. */
int pb_index_expand_term(pb_index* ix, const uint8_t* term, uint64_t term_len, uint8_t* out,
                         uint64_t cap, uint64_t* n_expansions, uint64_t* needed);
/* Per-term live occurrence count as the device computed it (tests). */
int pb_index_term_df_live(pb_index* ix, uint64_t* out, uint64_t cap);

/* How the posting columns are held in HBM.  The image above always carries u32 columns; at
 * pb_index_create the device copy becomes NARROW when, for every field,
 * (max tf + 1) << bits(max field length) <= 65536: one u16 code = tf << fl_bits | fl per field,
 * i.e. 4 + 2F bytes per row instead of 4 + 8F.  Results do not depend on the layout.
 * (Environment override for tests: PB_POSTING_LAYOUT=wide|narrow|auto.) */
typedef struct pb_device_layout {
  uint32_t narrow;              /* 1: u16 (tf, fl) codes, 0: u32 tf and field-length columns */
  uint32_t bytes_per_row;       /* 4 + 2F or 4 + 8F: the bytes the scoring kernel reads per posting row */
  uint32_t fl_bits[PB_MAX_FIELDS];
  uint64_t posting_bytes;       /* HBM held by the posting columns */
} pb_device_layout;
int pb_index_device_layout(pb_index* ix, pb_device_layout* out);

/* Diagnostic used by bench.py: read bandwidth (GB/s) of a plain streaming kernel (128-bit loads,
 * grid = 8 CTAs per SM) over a private buffer of `bytes`, averaged over `iters` passes after one
 * warm-up pass.  A buffer well below the 126 MB L2 measures the L2 -> SM fabric, a multi-GB buffer
 * measures HBM: the two ceilings the scoring kernel is compared with. */
int pb_device_read_bandwidth(int device, uint64_t bytes, uint32_t iters, double* gb_per_s);

/* ------------------------------------------------------------------------------------------
 * Queries (replaces Index::query, src/query.rs:21-106, for batches)
 * ---------------------------------------------------------------------------------------- */
typedef struct pb_query_batch_desc {
  uint64_t n_queries;
  const uint64_t* query_term_off; /* [n_queries + 1] indices into the term arrays; the terms of a
                                     query are ALL tokens the tokenizer produced, empty ones
                                     included (query.rs:32-35) */
  const uint64_t* term_byte_off;  /* [n_terms + 1] byte offsets into term_bytes */
  const uint8_t* term_bytes;      /* UTF-8 */
  uint32_t scorer;                /* PB_SCORER_* */
  double bm25_k1, bm25_b;         /* BM25 { bm25k1, bm25b }, bm25.rs:14-26 */
  const double* fields_boost;     /* [num_fields] (query.rs:26); finite values */
  uint32_t n_fields_boost;
  uint32_t top_k;                 /* <= PB_MAX_TOP_K */
} pb_query_batch_desc;

/* Per-query outputs.  Any pointer may be NULL (that output is skipped).
 * Digests (order independent, wrap-around sums over the result set):
 *   a(d)        = (u32)(d+1) * 0x9E3779B1                                          (32-bit, wraps)
 *   doc_digest  = sum over the result set of (u64)a(d) * a(d)                      (64-bit sum)
 *   score_digest= sum of (u64)y * y,  y = lo(s) ^ (u32)(hi(s)*0x85EBCA77) ^ a(d)   (64-bit sum)
 *                 (lo/hi = the two 32-bit halves of the f64 score's bit pattern)
 * top-k rows are ordered (score desc, doc ordinal asc) — the reference's comparison rule
 * (src/lib.rs:54-58) when ordinals follow key order. */
typedef struct pb_query_results {
  uint64_t* n_results;    /* [n_queries] size of the full result set (query.rs:97-100) */
  uint64_t* doc_digest;   /* [n_queries] */
  uint64_t* score_digest; /* [n_queries] */
  uint32_t* topk_n;       /* [n_queries] = min(top_k, n_results) */
  uint32_t* topk_doc;     /* [n_queries * top_k] doc ordinals */
  double* topk_score;     /* [n_queries * top_k] */
} pb_query_results;

/* One call, host buffers in / host buffers out (upload + kernels + download). */
int pb_query_batch(pb_index* ix, const pb_query_batch_desc* q, pb_query_results* out);

/* The same three stages separately: upload once, run (device-resident inputs), fetch. */
int pb_batch_create(pb_index* ix, const pb_query_batch_desc* q, pb_batch** out);
int pb_batch_run(pb_batch* b);
int pb_batch_fetch(pb_batch* b, pb_query_results* out);
void pb_batch_destroy(pb_batch* b);
/* DEVICE pointers of the batch's result buffers (same layout as pb_query_results), valid until
 * the batch is re-run or destroyed: lets a caller hand the top-k block straight to NCCL for the
 * multi-GPU gather without a host round trip. */
int pb_batch_device_results(pb_batch* b, pb_query_results* out_device_ptrs);

typedef struct pb_batch_stats {
  uint64_t n_queries, n_query_terms;
  uint64_t n_segments;        /* (query term, expanded term) posting lists walked */
  uint64_t rows_streamed;     /* posting rows read by the scoring kernel (direct + diverted) */
  uint64_t rows_streamed_direct; /* ... of which by the single launch over all single-list queries */
  uint64_t rows_scored;       /* rows whose doc is live = ScoreCalculator::score evaluations on
                                 de-duplicated rows ("scored postings") */
  uint64_t rows_diverted;     /* rows routed through the per-doc fold side path */
  uint64_t legacy_records;    /* ... of which through the sorted fallback (overflowing bins) */
  uint64_t results_emitted;   /* sum of n_results */
  uint64_t pointer_visits;    /* reference-equivalent DocumentPointer visits (sum of multiplicities) */
  uint32_t gpu_launches;      /* kernels launched by the last pb_batch_run */
  uint32_t side_rounds;       /* sub-batches of the side path */
  float ms_total;             /* CUDA-event time of the last pb_batch_run on its stream */
  float ms_descend, ms_plan, ms_score, ms_side, ms_finalize;
  uint32_t score_launches;    /* launches of the scoring kernel inside ms_score */
  float ms_gather;            /* the ncclAllGather of the result block (0 without pb_batch_set_gather); inside ms_total */
  /* CUDA-event time of each launch class of the multi-list side path, summed over its rounds (inside ms_side) */
  float ms_side_mark, ms_side_score, ms_side_fold;
  float ms_union;             /* the dense union kernel (ZeroToOne, union-heavy queries); inside ms_side */
  uint64_t rows_streamed_side;  /* posting rows read by the class-G scoring launches (inside ms_side_score) */
  uint64_t rows_streamed_union; /* posting rows read by the union kernel */
  uint64_t union_queries;       /* queries answered by the union kernel */
  uint64_t rows_streamed_compact; /* of rows_streamed_direct: rows read from the compact copy of the tiles (u16 doc offsets,
                                     2 + 2F bytes per row instead of 4 + 2F; built only with PB_POSTING_COMPACT=1) */
} pb_batch_stats;
int pb_batch_get_stats(const pb_batch* b, pb_batch_stats* out);
/* Stats of the last pb_query_batch / pb_query_full call on this index. */
int pb_index_last_stats(pb_index* ix, pb_batch_stats* out);

/* Full result sets (parity sampling; what Index::query returns): every (query, doc, score) of
 * every query of the batch, in no particular order.  *n_total is always set; returns
 * PB_ERR_CAPACITY when cap is too small (nothing useful written). */
int pb_query_full(pb_index* ix, const pb_query_batch_desc* q, uint64_t cap, uint32_t* out_query,
                  uint32_t* out_doc, double* out_score, uint64_t* n_total);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU.  Index::query is a pure function of &self (src/query.rs:21-22), so a batch
 * shards by QUERY: every GPU holds a replica of the image, rank r scores its own block of queries,
 * and the only exchange is ONE ncclAllGather of the packed per-query result blocks, issued by the
 * library on the batch's own stream right behind its last kernel.  NCCL is bound at run time
 * (libnccl.so.2 via dlopen); without it these entry points return PB_ERR_UNSUPPORTED.
 *
 * Two forms:
 *  (1) one PROCESS per GPU (torchrun / MPI / any launcher): rank 0 calls pb_comm_unique_id, the id
 *      reaches the other ranks over any host channel, every rank calls pb_comm_create, attaches the
 *      communicator to its batch with pb_batch_set_gather, and from then on pb_batch_run ends with
 *      the gather (inside ms_total); pb_batch_fetch_gathered returns all ranks' results.
 *  (2) one process, several devices: pb_group_create replicates an image on the listed devices
 *      (ncclCommInitAll) and pb_group_query_batch runs a whole batch across them.
 * Every rank must use the same slot_queries (>= its own n_queries) and top_k: rank r's query i
 * is global query r * slot_queries + i.  As with any collective, a rank that fails before the
 * gather leaves its peers waiting in it; pb_group_query_batch avoids that by gathering only after
 * every member has finished its kernels.
 * ---------------------------------------------------------------------------------------- */
#define PB_COMM_ID_BYTES 128u
typedef struct pb_comm pb_comm;   /* one rank of an NCCL communicator */
typedef struct pb_group pb_group; /* one process: an image replica + communicator rank per device */
int pb_comm_unique_id(uint8_t* id /* [PB_COMM_ID_BYTES] */);
int pb_comm_create(const uint8_t* id, int rank, int world, int device, pb_comm** out);
int pb_comm_info(const pb_comm* c, int* rank, int* world, int* nccl_version);
void pb_comm_destroy(pb_comm* c);

/* Attach (comm != NULL) or detach (NULL) the gather.  The communicator must outlive the batch. */
int pb_batch_set_gather(pb_batch* b, pb_comm* comm, uint64_t slot_queries);
/* Re-stage a new query batch into an existing pb_batch (keeps its stream, workspace and gather). */
int pb_batch_reload(pb_batch* b, const pb_query_batch_desc* q);
/* The stages of pb_batch_run separately: kernels only / enqueue the gather / wait for the stream. */
int pb_batch_run_local(pb_batch* b);
int pb_batch_gather(pb_batch* b);
int pb_batch_sync(pb_batch* b);
/* Results of ALL ranks in global query order: the first n_total of world * slot_queries queries. */
int pb_batch_fetch_gathered(pb_batch* b, uint64_t n_total, pb_query_results* out);
/* DEVICE view of the gathered blocks: rank r's packed block starts at block0 + r * block_bytes and
 * is laid out for slot_queries queries as [n_results u64][doc_digest u64][score_digest u64]
 * [topk_score f64 x k][topk_doc u32 x k][topk_n u32], each array slot_queries long. */
int pb_batch_device_gathered(pb_batch* b, const void** block0, uint64_t* block_bytes, uint64_t* slot_queries);

int pb_group_create(const pb_index_image* image, const int* devices, int n, pb_group** out);
int pb_group_size(const pb_group* g);
int pb_group_set_live_state(pb_group* g, const uint32_t* removed_ords, uint64_t n_removed, uint64_t n_live_docs,
                            const double* field_avg);
/* Queries are cut into contiguous blocks of ceil(n_queries / n) per member; `out` receives all
 * n_queries results in the caller's order (same layout as pb_query_batch). */
int pb_group_query_batch(pb_group* g, const pb_query_batch_desc* q, pb_query_results* out);
int pb_group_member_stats(pb_group* g, int member, pb_batch_stats* out);
void pb_group_destroy(pb_group* g);

/* Pinned host memory for the buffers of pb_query_batch (optional; any host memory works). */
void* pb_host_alloc(size_t bytes);
void pb_host_free(void* p);

const char* pb_last_error(void);
const char* pb_version(void);

/* Layout pins for foreign bindings that mirror these structs by hand (rust/src/lib.rs, capi.py): LP64, natural alignment. */
#if defined(__cplusplus)
#define PB_STATIC_ASSERT(c, m) static_assert(c, m)
#elif defined(__STDC_VERSION__) && __STDC_VERSION__ >= 201112L
#define PB_STATIC_ASSERT(c, m) _Static_assert(c, m)
#else
#define PB_STATIC_ASSERT(c, m)
#endif
PB_STATIC_ASSERT(sizeof(pb_index_image) == 248, "pb_index_image layout");
PB_STATIC_ASSERT(sizeof(pb_query_batch_desc) == 72, "pb_query_batch_desc layout");
PB_STATIC_ASSERT(sizeof(pb_query_results) == 48, "pb_query_results layout");
PB_STATIC_ASSERT(sizeof(pb_batch_stats) == 168, "pb_batch_stats layout");
PB_STATIC_ASSERT(sizeof(pb_builder_info) == 128, "pb_builder_info layout");
PB_STATIC_ASSERT(sizeof(pb_doc_tokens) == 32, "pb_doc_tokens layout");

#ifdef __cplusplus
}
#endif
#endif /* PROBLY_B200_H */
