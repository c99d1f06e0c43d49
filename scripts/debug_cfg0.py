"""Debug helper: cfg0_bench at full size, product vs oracle per query (both index-creation paths)."""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import oracle as orc
from probly_search_b200 import workload as W, Index, score, capi, DeviceBatch

name = sys.argv[1] if len(sys.argv) > 1 else "cfg0_bench"
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
cfg = W.CONFIGS[name]
wl = W.Workload(cfg)
o = orc.OracleIndex(cfg.n_fields); wl.build_into(o)
fq = wl.queries(nq)
exp = o.query_batch_flat(fq.query_term_off, fq.term_bytes, fq.term_byte_off, orc.BM25, cfg.boosts, 10, n_threads=8)
print("oracle: score_calls", exp["score_calls"], "results", int(exp["n_results"].sum()))

def report(tag, got, st):
    bad = np.nonzero(got.n_results != exp["n_results"])[0]
    print(f"[{tag}] rows_scored {st['rows_scored']} pointer_visits {st['pointer_visits']} results {st['results_emitted']} "
          f"segments {st['n_segments']} direct {st['rows_streamed_direct']} side {st['rows_streamed_side']} diverted {st['rows_diverted']} "
          f"rounds {st['side_rounds']}; queries with wrong n_results: {len(bad)}")
    for q in bad[:10]:
        print("    q", q, fq.terms_of(int(q)), "got", int(got.n_results[q]), "exp", int(exp["n_results"][q]))
    dd = np.nonzero((got.doc_digest != exp["doc_digest"]) | (got.score_digest != exp["score_digest"]))[0]
    print(f"    digest mismatches: {len(dd)}  first {dd[:10].tolist()}")

ix = Index(cfg.n_fields); wl.build_into(ix)
for n in (100, 300, nq):
    sub = fq.slice(0, n)
    b = DeviceBatch(ix, sub, score.bm25.new(), cfg.boosts, top_k=10)
    b.run(); got = b.fetch(); st = b.stats()
    e = {k: (v[:n] if hasattr(v, "__len__") else v) for k, v in exp.items()}
    bad = int((got.n_results != exp["n_results"][:n]).sum())
    print(f"device-flatten, first {n} queries: wrong n_results {bad}, rows_scored {st['rows_scored']}, results {st['results_emitted']} (oracle {int(exp['n_results'][:n].sum())})")
    b.close()
b = DeviceBatch(ix, fq, score.bm25.new(), cfg.boosts, top_k=10)
b.run(); report("device-flatten", b.fetch(), b.stats()); b.close()
# one query at a time through pb_query_batch
wrong = 0
for q in range(min(nq, 200)):
    r = ix.query_batch_flat(fq.slice(q, q + 1), score.bm25.new(), cfg.boosts, 10)
    if int(r.n_results[0]) != int(exp["n_results"][q]):
        wrong += 1
        if wrong <= 5: print("   single q", q, fq.terms_of(q), int(r.n_results[0]), int(exp["n_results"][q]))
print("single-query calls with wrong n_results among the first 200:", wrong)
# host-flattened image through pb_index_create (image file path)
path = "/dev/shm/pb_debug_cfg0.img"
ix.save_image(path)
ix2 = Index.load_image(path)
b = DeviceBatch(ix2, fq, score.bm25.new(), cfg.boosts, top_k=10)
b.run(); report("host-flatten", b.fetch(), b.stats()); b.close()
os.unlink(path)
