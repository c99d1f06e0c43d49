#!/usr/bin/env python
"""bench.py — the driver's benchmark contract for the probly-search query hot path.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
  python bench.py --impl reference ...                     (the CPU restatement, timed on host cores)

A "step" is one pass of the hot path over one batch of synthetic queries.  At N = 1 the workload
is BASELINE.json configs[1]: 1M-doc 2-field Zipfian corpus, 100k single-term BM25 queries.
At N > 1 every rank holds a replica of the index image and runs its own 100k-query batch (weak
scaling, queries are independent units); the only collective is the NCCL all-gather of the
per-query top-k blocks, straight from the library's device result buffers.

Prints ONE JSON line (rank 0).  `value` = scored postings / s with inputs resident in HBM;
`e2e` = the same metric through pb_query_batch with HOST buffers (H2D + kernels + D2H timed).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "scored-postings/sec"
UNIT = "postings/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2])); power.append(float(p[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


def build_product_index(cfg, n_docs, vocab, device):
    from probly_search_b200 import Index
    from probly_search_b200 import workload as W
    wl = W.Workload(cfg, n_docs=n_docs, vocab=vocab)
    ix = Index(cfg.n_fields, device=device)
    t = time.time()
    wl.build_into(ix)
    for d in wl.removed_ordinals():
        ix.remove_document(int(d))
    ix.sync_device()
    log(f"[bench] index built + uploaded in {time.time() - t:.1f}s: {ix.info().n_rows} rows")
    return wl, ix


def queries_for_rank(wl, n_queries, rank):
    """Rank r gets the r-th consecutive block of the config's query stream."""
    fq = wl.queries(n_queries * (rank + 1))
    return fq.slice(n_queries * rank, n_queries * (rank + 1)) if rank else fq


def cpu_arm(cfg, wl, n_docs, sample_fq, scorer_id, threads, target_s=15.0):
    """The oracle (CPU restatement of the reference) timed on the host cores, bounded sample."""
    from oracle import oracle as orc
    t = time.time()
    o = orc.OracleIndex(cfg.n_fields)
    wl.build_into(o)
    for d in wl.removed_ordinals():
        o.remove_document(int(d))
    log(f"[bench] oracle index built in {time.time() - t:.1f}s")

    def run(fq):
        return o.query_batch_flat(fq.query_term_off, fq.term_bytes, fq.term_byte_off, scorer_id, cfg.boosts, 10,
                                  n_threads=threads)
    probe_n = min(64, sample_fq.n_queries)
    r = run(sample_fq.slice(0, probe_n))
    per_q = max(r["seconds"] / probe_n, 1e-7)
    n = int(max(probe_n, min(sample_fq.n_queries, target_s / per_q)))
    return o, run, n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg1")
    ap.add_argument("--docs", type=int, default=None, help="override corpus size (debug only; invalidates the number)")
    ap.add_argument("--queries", type=int, default=None)
    ap.add_argument("--vocab", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--top-k", type=int, default=10)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 0)

    from probly_search_b200 import workload as W
    cfg = W.CONFIGS[args.config]
    n_docs = args.docs or cfg.n_docs
    vocab = args.vocab or cfg.vocab
    n_queries = args.queries or min(cfg.n_queries, 100_000)
    scorer_id = 0 if cfg.scorer == "bm25" else 1
    config = {"workload": f"{cfg.name}: {n_docs}-doc {cfg.n_fields}-field Zipfian corpus (V={vocab}), "
                          f"{n_queries} {'single-term' if cfg.query_mode == 0 else 'multi-term prefix'} "
                          f"{cfg.scorer} queries per GPU per step, top_k={args.top_k}",
              "boosts": list(cfg.boosts), "removed_fraction": cfg.removed_fraction,
              "l2": "inputs larger than L2 (index image >> 126 MB, every step streams it from HBM)",
              "parallelism": f"query-sharded x{world}, index replicated"}

    # ------------------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        wl = W.Workload(cfg, n_docs=n_docs, vocab=vocab)
        threads = os.cpu_count() or 1
        fq_all = wl.queries(min(n_queries, 20_000))
        o, run, n = cpu_arm(cfg, wl, n_docs, fq_all, scorer_id, threads, target_s=12.0)
        fq = fq_all.slice(0, n)
        for _ in range(warmup):
            run(fq)
        secs, ptr = 0.0, 0
        for _ in range(args.steps):
            r = run(fq)
            secs += r["seconds"]
            ptr = r["score_calls"]
        rows = int(r["n_results"].sum())     # single-term BM25, boosts > 0: results == scored de-duplicated rows
        if cfg.query_mode != 0 or any(b <= 0 for b in cfg.boosts):
            rows = ptr                        # otherwise report pointer visits
        val = rows * args.steps / secs
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": warmup, "ms_per_step": 1e3 * secs / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config,
                "queries_per_sec": fq.n_queries * args.steps / secs,
                "pointer_visits_per_sec": ptr * args.steps / secs,
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                                 "sample": f"first {fq.n_queries} queries of the workload per step, "
                                           f"oracle/probly_oracle.cpp (structure-faithful C++ restatement; the Rust "
                                           f"crate cannot be built here), {threads} threads"},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    # ------------------------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    from probly_search_b200 import DeviceBatch, capi, score

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    wl, ix = build_product_index(cfg, n_docs, vocab, local_rank)
    fq = queries_for_rank(wl, n_queries, rank)
    calc = score.bm25.new() if scorer_id == 0 else score.zero_to_one.new()
    k = args.top_k
    batch = DeviceBatch(ix, fq, calc, cfg.boosts, top_k=k)

    # NCCL gather of the per-query top-k blocks, straight from the library's device result buffers
    from probly_search_b200 import distributed as D
    gathered = None

    def step_device():
        nonlocal gathered
        batch.run()
        if world > 1:
            dp = batch.device_results()
            dev = f"cuda:{local_rank}"
            n = torch.as_tensor(D.DeviceArray(dp.topk_n, (n_queries,), "<i4"), device=dev)
            docs = torch.as_tensor(D.DeviceArray(dp.topk_doc, (n_queries, k), "<i4"), device=dev)
            scs = torch.as_tensor(D.DeviceArray(dp.topk_score, (n_queries, k), "<f8"), device=dev)
            gathered = D.gather_topk(n, docs, scs)      # NCCL over NVLink: the only collective of the path

    for _ in range(warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    dev_ms, score_ms, stats_acc = 0.0, 0.0, None
    for _ in range(args.steps):
        step_device()
        st = batch.stats()
        dev_ms += st["ms_total"]
        score_ms += st["ms_score"]
        stats_acc = st
    barrier()
    wall_s = time.perf_counter() - t0
    clocks = sampler.stop()
    st = stats_acc

    # max over ranks of the timed region (device-event time per step and wall clock)
    tvec = torch.tensor([dev_ms / 1e3, wall_s], dtype=torch.float64, device=f"cuda:{local_rank}")
    cnt = torch.tensor([st["rows_scored"], st["n_queries"], st["pointer_visits"], st["results_emitted"]],
                       dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(tvec, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    dev_s, wall_max = float(tvec[0]), float(tvec[1])
    rows_all, q_all, ptr_all, res_all = [float(x) for x in cnt]
    timed_s = wall_max if world > 1 else max(dev_s, 1e-9)   # N>1: the gather is outside the library's events
    value = rows_all * args.steps / timed_s

    # ---- e2e: host buffers through pb_query_batch (H2D + kernels + D2H inside the timed region)
    L = capi.lib()

    def pinned(arr):
        p = L.pb_host_alloc(arr.nbytes + 64)
        buf = (C.c_uint8 * (arr.nbytes + 64)).from_address(p)
        out = np.frombuffer(buf, dtype=arr.dtype, count=arr.size)
        out[:] = arr.ravel()
        return p, out
    p1, qoff = pinned(fq.query_term_off); p2, toff = pinned(fq.term_byte_off); p3, tbytes = pinned(fq.term_bytes)
    from probly_search_b200.index import BatchResults, FlatQueries
    fq_pinned = FlatQueries.__new__(FlatQueries)
    fq_pinned.query_term_off, fq_pinned.term_byte_off, fq_pinned.term_bytes = qoff, toff, tbytes
    res = BatchResults(n_queries, k)
    pins = []
    for name in ("n_results", "doc_digest", "score_digest", "topk_n", "topk_doc", "topk_score"):
        a = getattr(res, name)
        p, v = pinned(a)
        pins.append(p)
        setattr(res, name, v.reshape(a.shape))
    d, _keep = ix._desc(fq_pinned, calc, cfg.boosts, k)
    rs = res.c_struct()
    h2d = int(qoff.nbytes + toff.nbytes + tbytes.nbytes)
    d2h = int(sum(getattr(res, n).nbytes for n in ("n_results", "doc_digest", "score_digest", "topk_n", "topk_doc", "topk_score")))
    for _ in range(2):
        capi.check(L.pb_query_batch(ix._ix, C.byref(d), C.byref(rs)))
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(e2e_steps):
        capi.check(L.pb_query_batch(ix._ix, C.byref(d), C.byref(rs)))
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_rows = float(res.n_results.sum()) if False else st["rows_scored"]
    ev = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(ev, op=dist.ReduceOp.MAX)
    e2e_value = rows_all * e2e_steps / float(ev[0])
    for p in [p1, p2, p3] + pins:
        L.pb_host_free(p)

    # ---- roofline of the dominant kernel (the single launch over all single-list queries)
    F = cfg.n_fields
    peak, peak_src = hbm_peak()
    lay = ix.device_layout()
    bpr = lay["bytes_per_row"]                      # 4 + 2F (u16 codes) or 4 + 8F (u32 columns): DESIGN.md section 3
    algo_bytes = st["rows_streamed_direct"] * bpr
    launch_ms = st["ms_score"] / max(st["score_launches"], 1)
    achieved = algo_bytes / (launch_ms * 1e-3) / 1e9 if launch_ms > 0 else 0.0
    traffic = ncu_traffic()
    # the two measured ceilings of a streaming read on THIS box: L2 -> SM fabric and HBM
    ceilings = {}
    if rank == 0:
        for name, nbytes, iters in (("l2_read_gbs", 64 << 20, 50), ("hbm_read_gbs", 4 << 30, 5)):
            g = C.c_double(0.0)
            if L.pb_device_read_bandwidth(local_rank, nbytes, iters, C.byref(g)) == 0:
                ceilings[name] = g.value
    roofline = {"bound": "hbm", "kernel": f"pbk::score_kernel<F={F},{cfg.scorer},direct,{'narrow' if lay['narrow'] else 'wide'}>",
                "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "bytes_per_row": bpr, "posting_layout": "u16 (tf,fl) codes" if lay["narrow"] else "u32 columns",
                "algorithmic_bytes_per_launch": algo_bytes, "launch_ms": launch_ms,
                "traffic": (traffic or {}).get("dram_bytes_per_launch"),
                "traffic_source": (traffic or {}).get("source"),
                "share_of_step": st["ms_score"] / st["ms_total"] if st["ms_total"] else None,
                "measured_stream_ceilings": ceilings,
                "rows_per_sec_this_launch": st["rows_streamed_direct"] / (launch_ms * 1e-3) if launch_ms > 0 else None}

    # ---- CPU baseline (rank 0, N = 1 only): the oracle on the host cores, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            threads = os.cpu_count() or 1
            sample_all = fq.slice(0, min(fq.n_queries, 20_000))
            o, run, n = cpu_arm(cfg, wl, n_docs, sample_all, scorer_id, threads, target_s=15.0)
            sfq = sample_all.slice(0, n)
            r = run(sfq)
            sb = DeviceBatch(ix, sfq, calc, cfg.boosts, top_k=k)
            sb.run()
            srows = sb.stats()["rows_scored"]
            g = sb.fetch()
            parity = bool(np.array_equal(g.doc_digest, r["doc_digest"]) and np.array_equal(g.score_digest, r["score_digest"])
                          and np.array_equal(g.n_results, r["n_results"]))
            r1 = o.query_batch_flat(sfq.slice(0, max(1, n // 16)).query_term_off, sfq.term_bytes, sfq.term_byte_off,
                                    scorer_id, cfg.boosts, 10, n_threads=1)
            cpu = {"value": srows / r["seconds"], "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"first {n} queries of the same batch ({srows} de-duplicated rows, {r['score_calls']} "
                             f"reference pointer visits), oracle/probly_oracle.cpp on {threads} host threads",
                   "queries_per_sec": n / r["seconds"],
                   "single_thread_pointer_visits_per_sec": r1["score_calls"] / r1["seconds"],
                   "parity_with_gpu_on_sample": parity}
            sb.close()
        except Exception as e:  # the baseline is reported, never the thing measured
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": warmup, "ms_per_step": 1e3 * timed_s / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "queries_per_sec": q_all * args.steps / timed_s,
                "pointer_visits_per_sec": ptr_all * args.steps / timed_s,
                "wall_ms_per_step": 1e3 * wall_max / args.steps, "device_ms_per_step": 1e3 * dev_s / args.steps,
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": 1e3 * float(ev[0]) / e2e_steps,
                        "queries_per_sec": q_all * e2e_steps / float(ev[0])},
                "gpu_launches": int(st["gpu_launches"]) * args.steps,
                "stage_ms": {kk: st[kk] for kk in ("ms_descend", "ms_plan", "ms_score", "ms_side", "ms_finalize")},
                "rows": {"scored_per_step": rows_all, "streamed_direct": st["rows_streamed_direct"],
                         "diverted": st["rows_diverted"], "legacy_records": st.get("legacy_records"), "results": res_all, "side_rounds": st["side_rounds"]},
                "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
