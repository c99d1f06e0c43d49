#!/usr/bin/env python
"""Generates tests/golden/fullsize_<cfg>.npz: the CPU oracle's answers (result counts, doc-id digests,
score-bit digests, top-10) for the LEADING queries of BASELINE.json's full-size configurations, so
that GPU runs at those sizes (bench.py --config cfg2|cfg3|cfg4, tests/test_gpu_fullsize.py) can be
checked bit for bit without rebuilding a 10M-doc oracle index on the GPU box.

  python scripts/make_fullsize_goldens.py cfg1 cfg2 cfg3 cfg4      (CPU only; cfg3+cfg4 share one
                                                                   10M-doc oracle build, ~25 GB RAM)
The oracle is test infrastructure (oracle/probly_oracle.cpp); nothing in the product reads these files.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc                      # noqa: E402
from probly_search_b200 import workload as W          # noqa: E402

N_QUERIES = {"cfg1": 512, "cfg2": 400, "cfg3": 256, "cfg4": 256}
OUT = os.path.join(ROOT, "tests", "golden")


def run(o, cfg, wl, name, threads):
    n = N_QUERIES[name]
    fq = wl.queries(n, mode=cfg.query_mode)
    scorer = orc.BM25 if cfg.scorer == "bm25" else orc.ZERO_TO_ONE
    t = time.time()
    r = o.query_batch_flat(fq.query_term_off, fq.term_bytes, fq.term_byte_off, scorer, cfg.boosts, 10, n_threads=threads)
    print(f"[{name}] {n} queries, {r['score_calls']} pointer visits in {time.time() - t:.1f}s", flush=True)
    np.savez_compressed(os.path.join(OUT, f"fullsize_{name}.npz"), n_queries=n, n_docs=wl.n_docs, vocab=wl.vocab,
                        n_results=r["n_results"], doc_digest=r["doc_digest"], score_digest=r["score_digest"],
                        topk_n=r["topk_n"], topk_key=r["topk_key"], topk_score=r["topk_score"],
                        score_calls=r["score_calls"])


def main():
    names = sys.argv[1:] or ["cfg1", "cfg2", "cfg3", "cfg4"]
    threads = os.cpu_count() or 1
    groups = {}
    for nm in names:
        cfg = W.CONFIGS[nm]
        groups.setdefault((cfg.cfg, cfg.n_docs, cfg.vocab), []).append(nm)
    for key, nms in groups.items():
        nms.sort(key=lambda nm: W.CONFIGS[nm].removed_fraction)      # the un-removed configs first
        cfg0 = W.CONFIGS[nms[0]]
        wl = W.Workload(cfg0)
        o = orc.OracleIndex(cfg0.n_fields)
        t = time.time()
        wl.build_into(o)
        print(f"oracle index for {nms}: {wl.n_docs} docs built in {time.time() - t:.1f}s", flush=True)
        removed_done = False
        for nm in nms:
            cfg = W.CONFIGS[nm]
            wln = W.Workload(cfg)
            if cfg.removed_fraction > 0 and not removed_done:
                for d in wln.removed_ordinals():
                    o.remove_document(int(d))
                removed_done = True
            run(o, cfg, wln, nm, threads)


if __name__ == "__main__":
    main()
