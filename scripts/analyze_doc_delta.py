"""SURVEY §8(f)-4, evaluated before building it: how many bytes would delta-coded doc ordinals save?

For every 128-row tile of the term-major image: does `doc - min(doc of the tile)` fit 16 bits?  Reported per tile
(= per stored byte) and weighted by how often a tile is streamed when single-term queries are Zipf-drawn (cfg 1/3/4:
the tile's term is drawn with probability ~ 1 / rank, so dense lists dominate the stream).

  python scripts/analyze_doc_delta.py cfg1            # 1 M docs, seconds
  python scripts/analyze_doc_delta.py cfg3 2000000    # the cfg 3 generator at 2 M docs (10 M needs ~40 GB of host RAM)
"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from probly_search_b200 import Index, workload as W

name = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
cfg = W.CONFIGS[name]
n_docs = int(sys.argv[2]) if len(sys.argv) > 2 else cfg.n_docs
wl = W.Workload(cfg, n_docs=n_docs)
ix = Index(cfg.n_fields)
wl.build_into(ix)
im = ix.flatten()
F, T = cfg.n_fields, 128
words = T * (1 + 2 * F)
n_rows, n_terms = int(im.n_rows), int(im.n_terms)
n_tiles = (n_rows + T - 1) // T
blocks = np.ctypeslib.as_array(im.post_blocks, shape=(int(im.n_rows_padded) // T * words,)).reshape(-1, words)
doc = blocks[:n_tiles, :T].astype(np.int64)
row_begin = np.ctypeslib.as_array(im.term_row_begin, shape=(n_terms + 1,)).astype(np.int64)
rows = np.arange(n_tiles * T).reshape(n_tiles, T)
valid = rows < n_rows
lo = np.where(valid, doc, 1 << 40).min(axis=1)
hi = np.where(valid, doc, -1).max(axis=1)
fits = (hi - lo) < 65536
# stream weight of a row = probability that its term is the query: the generator draws a vocabulary word by Zipf rank;
# approximate the rank of a term by the rank of its list length (exact for the head, where it matters)
df = np.diff(row_begin)
order = np.argsort(-df, kind="stable")
rank = np.empty(n_terms, dtype=np.int64); rank[order] = np.arange(1, n_terms + 1)
p_term = 1.0 / rank
term_of_row_tile = np.searchsorted(row_begin, np.arange(n_tiles) * T, side="right") - 1   # term of the tile's first row
w = p_term[np.clip(term_of_row_tile, 0, n_terms - 1)]
print(f"{name} at {n_docs} docs: {n_rows} rows, {n_terms} terms, {n_tiles} tiles")
print(f"  tiles whose doc span fits u16: {fits.mean():.3f} of the stored tiles, {np.average(fits, weights=w):.3f} of the streamed tiles")
for F_ in (F,):
    now = 4 + 2 * F_
    new = 2 + 2 * F_
    s_store = fits.mean(); s_stream = np.average(fits, weights=w)
    print(f"  bytes/row {now} -> {now - 2 * s_store:.2f} stored, {now - 2 * s_stream:.2f} streamed (all-fit bound {new}); "
          f"+4 B/tile for the base, +1 branch per tile for the format")
