"""Builds the cfg1 index on the device and applies one removal (for `ncu -k regex:live_df|row_dead` timing of the live-state kernels)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from probly_search_b200 import workload as W, Index
cfg = W.CONFIGS["cfg1"]
wl = W.Workload(cfg)
ix = Index(cfg.n_fields)
wl.build_into(ix)
t = time.time(); ix.sync_device(); print("sync_device (create)", round(time.time() - t, 2), "s")
ix.remove_document(12345)
t = time.time(); ix.sync_device(); print("sync_device (one removal -> pb_index_set_live_state)", round((time.time() - t) * 1e3, 2), "ms")
