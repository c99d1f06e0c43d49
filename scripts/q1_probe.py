"""Q = 1 through pb_query_batch, cfg0-sized corpus: wall-clock split (host call vs the library's own CUDA-event time) and,
under `ncu --metrics gpu__time_duration.sum --launch-skip 600 --launch-count 42`, the kernels of a few such queries."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from probly_search_b200 import workload as W, Index, score

name = sys.argv[1] if len(sys.argv) > 1 else "cfg0"
cfg = W.CONFIGS[name]
n_docs = int(sys.argv[2]) if len(sys.argv) > 2 else cfg.n_docs
wl = W.Workload(cfg, n_docs=n_docs)
ix = Index(cfg.n_fields)
wl.build_into(ix)
fq = wl.queries(200)
calc = score.bm25.new()
ones = [fq.slice(q, q + 1) for q in range(200)]
for q in range(150):
    ix.query_batch_flat(ones[q], calc, cfg.boosts, 10)
ts, dev = [], []
for q in range(200):
    t = time.perf_counter()
    ix.query_batch_flat(ones[q], calc, cfg.boosts, 10)
    ts.append(time.perf_counter() - t)
    dev.append(ix.last_stats()["ms_total"] * 1e3)
ts = np.asarray(ts) * 1e6
print(f"{name}: wall p50 {np.percentile(ts, 50):.1f} us, mean {ts.mean():.1f}; library CUDA-event time (first kernel -> last) p50 "
      f"{np.percentile(dev, 50):.1f} us, mean {np.mean(dev):.1f}; launches {ix.last_stats()['gpu_launches']}")
