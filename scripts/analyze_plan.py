"""Host-side look at how a config's query batch splits into classes (single list / primary scheme /
exact scheme) and how many rows each class streams.  No GPU needed."""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probly_search_b200 import Index, workload as W

cfgname = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
cfg = W.CONFIGS[cfgname]
wl = W.Workload(cfg)
ix = Index(cfg.n_fields)
wl.build_into(ix)
im = ix.flatten()
nn, ne, nt = int(im.n_nodes), int(im.n_edges), int(im.n_terms)
A = np.ctypeslib.as_array
eb = A(im.node_edge_begin, shape=(nn + 1,)); tlo = A(im.node_term_lo, shape=(nn,)); thi = A(im.node_term_hi, shape=(nn,))
ec = A(im.edge_char, shape=(ne,)); ech = A(im.edge_child, shape=(ne,)); trb = A(im.term_row_begin, shape=(nt + 1,))
rows = np.diff(trb.astype(np.int64))
fq = wl.queries(nq)
cls = {"S": [0, 0], "P": [0, 0, 0], "X": [0, 0]}
nseg = 0
sec_hist = []
for q in range(nq):
    tot, big, n = 0, 0, 0
    for t in fq.terms_of(q):
        node = 0; ok = len(t) > 0
        for ch in t:
            a, b = eb[node], eb[node + 1]
            i = a + np.searchsorted(ec[a:b], ord(ch))
            if i < b and ec[i] == ord(ch): node = ech[i]
            else: ok = False; break
        if not ok: continue
        r = rows[tlo[node]:thi[node]]
        r = r[r > 0]
        n += len(r); tot += int(r.sum()); big = max(big, int(r.max()) if len(r) else 0)
    nseg += n
    if n == 1: cls["S"][0] += 1; cls["S"][1] += tot
    elif n >= 2:
        sec = tot - big
        if sec * 16 >= tot: cls["X"][0] += 1; cls["X"][1] += tot
        else: cls["P"][0] += 1; cls["P"][1] += tot; cls["P"][2] += sec; sec_hist.append((sec, big))
print("segments", nseg)
print("S queries %d rows %.3e" % tuple(cls["S"]))
print("P queries %d rows %.3e secondary rows %.3e" % tuple(cls["P"]))
print("X queries %d rows %.3e" % tuple(cls["X"]))
sh = np.array(sec_hist)
if len(sh):
    # primary rows of queries by secondary size buckets
    for lo, hi in ((0, 100), (100, 1000), (1000, 10000), (10000, 10**9)):
        m = (sh[:, 0] >= lo) & (sh[:, 0] < hi)
        print(f"P queries with secondary rows in [{lo},{hi}): {m.sum()} queries, primary rows {sh[m,1].sum():.3e}")
