import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probly_search_b200 import DeviceBatch, Index, score, workload as W
cfg = W.CONFIGS["cfg1"]
wl = W.Workload(cfg)
ix = Index(2)
wl.build_into(ix)
fq = wl.queries(300, mode=1)
b = DeviceBatch(ix, fq, score.zero_to_one.new(), cfg.boosts, top_k=10)
b.run(); got = b.fetch(); st = b.stats()
print("n_results[0]", int(got.n_results[0]), st["legacy_records"], st["rows_diverted"])
