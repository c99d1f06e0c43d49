#!/usr/bin/env python
"""bench.py — the driver's benchmark contract for the probly-search query hot path.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
  python bench.py --impl reference ...                     (the CPU restatement, timed on host cores)
  python bench.py --config cfg2|cfg3|cfg4[,cfgX...]        (the other BASELINE.json configs; one JSON line each)

A "step" is one pass of the hot path over one batch of synthetic queries.  Default workload =
BASELINE.json configs[1]: 1M-doc 2-field Zipfian corpus, 100k single-term BM25 queries per GPU
(weak scaling: every rank holds a replica of the index image and runs its own 100k-query block of the
query stream).  cfg3 / cfg4 are the 8-GPU configurations: a 10M-doc image replicated per GPU and ONE
batch of 1M queries sharded by query over the ranks (strong scaling).

The only collective of the path is one ncclAllGather of the packed per-query result blocks, issued by
the LIBRARY on the batch's own stream behind its last kernel (pb_batch_set_gather): it is inside the
timed region of both `value` and `e2e`.  torch.distributed is used for the rendezvous only.

Prints ONE JSON line per config (rank 0).  `value` = scored postings / s with inputs resident in HBM;
`e2e` = the same metric with HOST buffers (H2D of the queries + kernels + gather + D2H of the results).
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
L2_BYTES = 126 << 20


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


def ncu_traffic(cfg_name: str, kernel_class: str, queries_per_gpu: int):
    """dram bytes per launch of the dominant kernel from a committed ncu capture of THIS config, launch class and
    batch size per GPU (profiles/roofline_traffic.json); anything else is not this launch's traffic -> None."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        e = d.get(cfg_name)
        if e and e.get("class") == kernel_class and e.get("queries_per_gpu", queries_per_gpu) == queries_per_gpu:
            return e
    except Exception:
        pass
    return None


def golden(cfg_name: str, n_docs: int, vocab: int):
    """The oracle's answers for the leading queries of a full-size config (scripts/make_fullsize_goldens.py)."""
    p = os.path.join(ROOT, "tests", "golden", f"fullsize_{cfg_name}.npz")
    if not os.path.exists(p):
        return None
    g = np.load(p)
    if int(g["n_docs"]) != n_docs or int(g["vocab"]) != vocab:
        return None
    return g


def check_against(res, exp, n) -> bool:
    ok = (np.array_equal(res.n_results[:n], exp["n_results"][:n]) and np.array_equal(res.doc_digest[:n], exp["doc_digest"][:n])
          and np.array_equal(res.score_digest[:n], exp["score_digest"][:n]) and np.array_equal(res.topk_n[:n], exp["topk_n"][:n]))
    for q in range(n if ok else 0):
        m = int(res.topk_n[q])
        ok = ok and np.array_equal(res.topk_doc[q, :m], np.asarray(exp["topk_key"][q, :m]).astype(np.uint32)) \
            and np.array_equal(res.topk_score[q, :m], exp["topk_score"][q, :m])
    return bool(ok)


def cpu_arm(cfg, wl, sample_fq, scorer_id, threads, target_s=15.0):
    """The oracle (CPU restatement of the reference) timed on the host cores, bounded sample."""
    from oracle import oracle as orc
    t = time.time()
    o = orc.OracleIndex(cfg.n_fields)
    wl.build_into(o)
    for d in wl.removed_ordinals():
        o.remove_document(int(d))
    log(f"[bench] oracle index built in {time.time() - t:.1f}s")

    def run(fq, n_threads=threads):
        return o.query_batch_flat(fq.query_term_off, fq.term_bytes, fq.term_byte_off, scorer_id, cfg.boosts, 10,
                                  n_threads=n_threads)
    probe_n = min(64, sample_fq.n_queries)
    r = run(sample_fq.slice(0, probe_n))
    per_q = max(r["seconds"] / probe_n, 1e-7)
    n = int(max(probe_n, min(sample_fq.n_queries, target_s / per_q)))
    return o, run, n


def workload_text(cfg, n_docs, vocab, n_queries, per, top_k):
    kind = "single-term" if cfg.query_mode == 0 else "multi-term prefix"
    return (f"{cfg.name}: {n_docs}-doc {cfg.n_fields}-field Zipfian corpus (V={vocab}), "
            f"{n_queries} {kind} {cfg.scorer} queries {per}, top_k={top_k}")


def reference_arm(args, cfg, n_docs, vocab, warmup):
    from probly_search_b200 import workload as W
    scorer_id = 0 if cfg.scorer == "bm25" else 1
    n_queries = args.queries or min(cfg.n_queries, 100_000)
    wl = W.Workload(cfg, n_docs=n_docs, vocab=vocab)
    threads = os.cpu_count() or 1
    fq_all = wl.queries(min(n_queries, 20_000))
    o, run, n = cpu_arm(cfg, wl, fq_all, scorer_id, threads, target_s=12.0)
    fq = fq_all.slice(0, n)
    for _ in range(warmup):
        run(fq)
    secs, ptr = 0.0, 0
    for _ in range(args.steps):
        r = run(fq)
        secs += r["seconds"]
        ptr = r["score_calls"]
    rows = int(r["n_results"].sum())     # single-term BM25, boosts > 0: results == scored de-duplicated rows
    unit_note = "results == de-duplicated scored rows for single-term queries"
    if cfg.query_mode != 0 or any(b <= 0 for b in cfg.boosts):
        rows = ptr                        # multi-list queries: the reference's own unit of work, score() calls
        unit_note = "ScoreCalculator::score calls (pointer visits) — the GPU arm counts de-duplicated rows, which is fewer"
    val = rows * args.steps / secs
    config = {"workload": workload_text(cfg, n_docs, vocab, n_queries, "per GPU per step", args.top_k),
              "boosts": list(cfg.boosts), "removed_fraction": cfg.removed_fraction,
              "parallelism": f"query-sharded x{args.gpus}, index replicated"}
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": warmup, "ms_per_step": 1e3 * secs / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config,
            "queries_per_sec": fq.n_queries * args.steps / secs,
            "pointer_visits_per_sec": ptr * args.steps / secs,
            "unit_note": unit_note,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"first {fq.n_queries} queries of the workload per step, "
                                       f"oracle/probly_oracle.cpp (structure-faithful C++ restatement; the Rust "
                                       f"crate cannot be built here), {threads} threads"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


class Pinned:
    """numpy views over pb_host_alloc memory (freed on close)."""

    def __init__(self, L):
        self.L, self.ptrs = L, []

    def like(self, arr):
        p = self.L.pb_host_alloc(arr.nbytes + 64)
        self.ptrs.append(p)
        buf = (C.c_uint8 * (arr.nbytes + 64)).from_address(p)
        out = np.frombuffer(buf, dtype=arr.dtype, count=arr.size).reshape(arr.shape)
        out[...] = arr
        return out

    def close(self):
        for p in self.ptrs:
            self.L.pb_host_free(p)
        self.ptrs = []


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg1", help="cfg1 (default), cfg2, cfg3, cfg4, cfg0; comma-separated list = one line each")
    ap.add_argument("--docs", type=int, default=None, help="override corpus size (debug only; invalidates the number)")
    ap.add_argument("--queries", type=int, default=None)
    ap.add_argument("--vocab", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-baseline", action="store_true", help="force the CPU baseline (10M-doc configs skip it by default)")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--top-k", type=int, default=10)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 0)

    from probly_search_b200 import workload as W
    names = [c.strip() for c in args.config.split(",") if c.strip()]

    if args.impl == "reference":
        if rank != 0:
            return
        cfg = W.CONFIGS[names[0]]
        reference_arm(args, cfg, args.docs or cfg.n_docs, args.vocab or cfg.vocab, warmup)
        return

    import torch
    import torch.distributed as dist
    from probly_search_b200 import capi
    from probly_search_b200 import distributed as D

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = D.Comm.from_torch(local_rank)          # the library's own communicator (rendezvous over torch)
        log(f"[bench] rank {rank}/{world}: library NCCL communicator up (NCCL {comm.nccl_version()})")

    ctx = {"args": args, "rank": rank, "local_rank": local_rank, "world": world, "warmup": warmup, "comm": comm,
           "torch": torch, "dist": dist, "index_cache": {}}
    for name in names:
        run_config(ctx, W.CONFIGS[name])
    for ent in ctx["index_cache"].values():
        ent[1].close()
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


def get_index(ctx, cfg, n_docs, vocab):
    """One image per corpus, replicated on every rank.  N > 1: rank 0 builds it with the host builder and
    writes the image file once (pb_image_save); the other ranks serve it from the file (pb_image_load)."""
    from probly_search_b200 import Index
    from probly_search_b200 import workload as W
    torch, dist = ctx["torch"], ctx["dist"]
    rank, world, dev = ctx["rank"], ctx["world"], ctx["local_rank"]
    key = (cfg.cfg, n_docs, vocab, cfg.n_fields, bool(getattr(cfg, 'bench_shape', False)))   # cfg3 / cfg4 share one corpus
    wl = W.Workload(cfg, n_docs=n_docs, vocab=vocab)
    if key in ctx["index_cache"]:
        _, ix, state = ctx["index_cache"][key]
    else:
        t = time.time()
        shm = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
        path = os.path.join(shm, f"pb_bench_{os.environ.get('MASTER_PORT', '0')}_{cfg.name}_{n_docs}_{vocab}.img")
        if rank == 0:
            ix = Index(cfg.n_fields, device=dev)
            wl.build_into(ix)
            if world > 1:
                ix.save_image(path)
        if world > 1:
            dist.barrier()
            if rank != 0:
                ix = Index.load_image(path, device=dev)
        ix.sync_device()
        if world > 1:
            dist.barrier()
            if rank == 0:
                os.unlink(path)
        state = {"removed": 0.0}
        ctx["index_cache"][key] = (wl, ix, state)
        log(f"[bench] rank {rank}: index image resident after {time.time() - t:.1f}s "
            f"({ix.device_layout()['posting_bytes'] / 1e6:.0f} MB of posting columns)")
    # live state of this config (removed-not-vacuumed docs): rank 0's builder is the source of truth
    if cfg.removed_fraction != state["removed"]:
        if state["removed"] != 0.0:
            raise SystemExit("bench.py: order the configs so that the un-removed corpus comes first")
        t = time.time()
        if rank == 0:
            for d in wl.removed_ordinals():
                ix.remove_document(int(d))
            ix.sync_device()                           # pb_index_set_live_state: no re-flatten, no re-upload
            ords, n_live, avg = ix.live_state()
            hdr = torch.tensor([len(ords), n_live] + [0] * 0, dtype=torch.int64, device=f"cuda:{dev}")
            favg = torch.tensor(avg + [0.0] * (4 - len(avg)), dtype=torch.float64, device=f"cuda:{dev}")
        else:
            hdr = torch.zeros(2, dtype=torch.int64, device=f"cuda:{dev}")
            favg = torch.zeros(4, dtype=torch.float64, device=f"cuda:{dev}")
        if world > 1:
            dist.broadcast(hdr, src=0)
            dist.broadcast(favg, src=0)
            n_rm = int(hdr[0])
            t_ords = torch.from_numpy(ords.astype(np.int64)).to(f"cuda:{dev}") if rank == 0 else \
                torch.zeros(n_rm, dtype=torch.int64, device=f"cuda:{dev}")
            dist.broadcast(t_ords, src=0)
            if rank != 0:
                ix.set_live_state(t_ords.cpu().numpy().astype(np.uint32), int(hdr[1]), [float(x) for x in favg[: cfg.n_fields]])
        state["removed"] = cfg.removed_fraction
        log(f"[bench] rank {rank}: live state ({cfg.removed_fraction:.0%} removed, pre-vacuum) applied in {time.time() - t:.1f}s")
    return wl, ix


def run_config(ctx, cfg):
    from probly_search_b200 import DeviceBatch, capi, score
    from probly_search_b200 import distributed as D
    from probly_search_b200.index import BatchResults, FlatQueries
    args, torch, dist = ctx["args"], ctx["torch"], ctx["dist"]
    rank, local_rank, world, warmup, comm = ctx["rank"], ctx["local_rank"], ctx["world"], ctx["warmup"], ctx["comm"]
    dev = f"cuda:{local_rank}"
    n_docs = args.docs or cfg.n_docs
    vocab = args.vocab or cfg.vocab
    scorer_id = 0 if cfg.scorer == "bm25" else 1
    k = args.top_k
    sharded = cfg.cfg >= 3                     # cfg3 / cfg4: ONE batch sharded by query (strong); else one block per rank (weak)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    wl, ix = get_index(ctx, cfg, n_docs, vocab)
    if sharded:
        total_q = args.queries or cfg.n_queries
        lo, hi, slot = D.slot_block(total_q, rank, world)
        fq = wl.queries(hi).slice(lo, hi)
        per = f"in ONE batch sharded by query over {world} GPU(s)"
        wtext = workload_text(cfg, n_docs, vocab, total_q, per, k)
    else:
        n_queries = args.queries or min(cfg.n_queries, 100_000)
        lo, hi, slot = n_queries * rank, n_queries * (rank + 1), n_queries
        fq = wl.queries(hi).slice(lo, hi) if rank else wl.queries(hi)
        total_q = n_queries * world
        wtext = workload_text(cfg, n_docs, vocab, n_queries, "per GPU per step", k)
    nq_local = fq.n_queries
    lay = ix.device_layout()
    config = {"workload": wtext, "boosts": list(cfg.boosts), "removed_fraction": cfg.removed_fraction,
              "l2": f"flushed between steps (256 MB memset); posting image = {lay['posting_bytes'] / 1e6:.0f} MB vs 126 MB of L2 "
                    f"— inside a step Zipf-drawn queries re-read hot lists, so L2 hits are part of the workload (see roofline)",
              "parallelism": f"query-sharded x{world}, index replicated",
              "collective": "one ncclAllGather of the packed result block per step, issued by the library on the batch stream"
                            if world > 1 else "none (1 GPU)"}

    calc = score.bm25.new() if scorer_id == 0 else score.zero_to_one.new()
    batch = DeviceBatch(ix, fq, calc, cfg.boosts, top_k=k)
    if comm is not None:
        batch.set_gather(comm, slot)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def l2_flush():
        flush.zero_()
        torch.cuda.synchronize()

    for _ in range(warmup):
        l2_flush()
        batch.run()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    dev_ms = 0.0
    acc = {}
    for _ in range(args.steps):
        l2_flush()
        batch.run()                      # kernels + (N > 1) the gather, all on the batch's stream
        st = batch.stats()
        dev_ms += st["ms_total"]
        for kk in ("ms_descend", "ms_plan", "ms_score", "ms_side", "ms_finalize", "ms_gather", "ms_side_mark",
                   "ms_side_score", "ms_side_fold", "ms_union"):
            acc[kk] = acc.get(kk, 0.0) + st[kk] / args.steps
    barrier()
    wall_s = time.perf_counter() - t0
    clocks = sampler.stop()

    # max over ranks of the timed region; sums of the units processed
    tvec = torch.tensor([dev_ms / 1e3, wall_s], dtype=torch.float64, device=dev)
    cnt = torch.tensor([st["rows_scored"], st["n_queries"], st["pointer_visits"], st["results_emitted"]],
                       dtype=torch.float64, device=dev)
    per_rank_ms = None
    if world > 1:
        mine = torch.tensor([dev_ms / args.steps, acc["ms_score"], acc["ms_side"], acc["ms_gather"]], dtype=torch.float64, device=dev)
        allr = torch.zeros(world * 4, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, mine)
        per_rank_ms = [[round(float(x), 3) for x in allr[4 * r: 4 * r + 4]] for r in range(world)]
        dist.all_reduce(tvec, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    dev_s, wall_max = float(tvec[0]), float(tvec[1])
    rows_all, q_all, ptr_all, res_all = [float(x) for x in cnt]
    timed_s = max(dev_s, 1e-9)           # CUDA events on the batch stream: kernels + gather
    value = rows_all * args.steps / timed_s

    # ---- parity: against the committed oracle answers for the leading queries (rank 0's block starts at query 0),
    #      and rank 0's gathered copy of every rank's block against that rank's own results
    local = batch.fetch()
    parity = {}
    g = golden(cfg.name, n_docs, vocab) if rank == 0 else None
    if g is not None:
        n = min(int(g["n_queries"]), nq_local)
        parity["golden_queries"] = n
        parity["golden_ok"] = check_against(local, g, n)
    if world > 1:
        def csum(r, a, b):
            return int(np.bitwise_xor.reduce(r.doc_digest[a:b]) ^ (np.bitwise_xor.reduce(r.score_digest[a:b]) << np.uint64(1))
                       ^ np.uint64(int(r.n_results[a:b].sum()) & 0xFFFFFFFF)) & 0x7FFFFFFFFFFFFFFF
        mine = torch.tensor([csum(local, 0, nq_local)], dtype=torch.int64, device=dev)
        allc = torch.zeros(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allc, mine)
        if rank == 0:
            gathered = batch.fetch_gathered(world * slot)
            ok = True
            for r in range(world):
                a = r * slot
                b = a + (slot if not sharded else D.slot_block(total_q, r, world)[1] - D.slot_block(total_q, r, world)[0])
                ok = ok and csum(gathered, a, b) == int(allc[r])
            parity["gathered_blocks_match_ranks"] = bool(ok)

    # ---- e2e: host buffers -> H2D (reload) -> kernels -> gather -> D2H of this rank's results, per step
    L = capi.lib()
    pin = Pinned(L)
    fq_pinned = FlatQueries.__new__(FlatQueries)
    fq_pinned.query_term_off = pin.like(fq.query_term_off)
    fq_pinned.term_byte_off = pin.like(fq.term_byte_off)
    fq_pinned.term_bytes = pin.like(fq.term_bytes)
    res = BatchResults(nq_local, k)
    for name in ("n_results", "doc_digest", "score_digest", "topk_n", "topk_doc", "topk_score"):
        setattr(res, name, pin.like(getattr(res, name)))
    rs = res.c_struct()
    h2d = int(fq_pinned.query_term_off.nbytes + fq_pinned.term_byte_off.nbytes + fq_pinned.term_bytes.nbytes)
    d2h = int(sum(getattr(res, n).nbytes for n in ("n_results", "doc_digest", "score_digest", "topk_n", "topk_doc", "topk_score")))

    def e2e_step():
        batch.reload(fq_pinned)                                        # H2D of the query batch
        batch.run()                                                    # kernels (+ gather)
        capi.check(L.pb_batch_fetch(batch._handle(), C.byref(rs)))     # D2H of the per-query results

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        l2_flush()
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    ev = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ev, op=dist.ReduceOp.MAX)
    e2e_value = rows_all * args.steps / float(ev[0])
    parity["e2e_equals_device_run"] = bool(np.array_equal(res.doc_digest, local.doc_digest) and
                                           np.array_equal(res.score_digest, local.score_digest))

    # ---- roofline: every launch class is event-timed separately; the dominant one (by time) is reported
    F = cfg.n_fields
    peak, peak_src = hbm_peak()
    bpr = lay["bytes_per_row"]                      # 4 + 2F (u16 codes) or 4 + 8F (u32 columns): DESIGN.md section 3
    bpr_survey = 4 + 8 * F                          # SURVEY §8(d)'s per-row figure (u32 columns)
    layout_name = "narrow" if lay["narrow"] else "wide"
    classes = {
        "direct": {"kernel": f"pbk::score_kernel<F={F},{cfg.scorer},direct,{layout_name}>", "ms": acc["ms_score"],
                   "rows": st["rows_streamed_direct"], "bytes_per_row": bpr, "launches": st["score_launches"]},
        "side_score": {"kernel": f"pbk::score_kernel<F={F},{cfg.scorer},divert,{layout_name}>", "ms": acc["ms_side_score"],
                       "rows": st["rows_streamed_side"], "bytes_per_row": bpr, "launches": st["side_rounds"]},
        "union": {"kernel": f"pbk::union_warp_kernel<F={F}> (PB_UNION_KERNEL=cta: pbk::union_kernel)", "ms": acc["ms_union"], "rows": st["rows_streamed_union"],
                  "bytes_per_row": 8 + 2 * F, "launches": 1 if st["union_queries"] else 0},
        "side_mark": {"kernel": f"pbk::mark_kernel<F={F}>", "ms": acc["ms_side_mark"], "rows": None, "bytes_per_row": None,
                      "launches": st["side_rounds"]},
        "side_fold": {"kernel": f"pbk::binfold_kernel<F={F},{cfg.scorer}>", "ms": acc["ms_side_fold"], "rows": None,
                      "bytes_per_row": None, "launches": st["side_rounds"]},
    }
    # rows of the single-list launch that streamed from the compact copy (u16 doc offsets: 2 B/row fewer)
    rows_compact = int(st.get("rows_streamed_compact", 0))
    if rows_compact and classes["direct"]["rows"]:
        d = classes["direct"]
        d["rows_compact"] = rows_compact
        d["bytes_per_row"] = (d["rows"] * bpr - 2 * rows_compact) / d["rows"]
        d["kernel"] += f" ({rows_compact / d['rows']:.1%} of the rows from the compact tiles, {bpr - 2} B/row)"
    for c in classes.values():
        c["share_of_step"] = c["ms"] / (dev_ms / args.steps) if dev_ms else None
        if c["rows"] and c["ms"] > 0:
            c["rows_per_sec"] = c["rows"] / (c["ms"] * 1e-3)
            c["layout_gbs"] = c["rows"] * c["bytes_per_row"] / (c["ms"] * 1e-3) / 1e9
    dom_name = max((n for n in classes if classes[n]["rows"]), key=lambda n: classes[n]["ms"], default="direct")
    dom = classes[dom_name]
    launch_ms = dom["ms"] / max(dom["launches"], 1)
    algo_bytes = (dom["rows"] or 0) * dom["bytes_per_row"] / max(dom["launches"], 1)
    achieved = algo_bytes / (launch_ms * 1e-3) / 1e9 if launch_ms > 0 else 0.0
    ceilings = {}
    if rank == 0:
        for name, nbytes, iters in (("l2_read_gbs", 64 << 20, 50), ("hbm_read_gbs", 4 << 30, 5)):
            gb = C.c_double(0.0)
            if L.pb_device_read_bandwidth(local_rank, nbytes, iters, C.byref(gb)) == 0:
                ceilings[name] = gb.value
    tr = ncu_traffic(cfg.name, dom_name, nq_local)
    image_fits_l2 = lay["posting_bytes"] < 2 * L2_BYTES
    roofline = {"bound": "hbm", "kernel": dom["kernel"], "class": dom_name,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "accounting": f"rows streamed by this launch class x {dom['bytes_per_row']:.3g} B/row of the device layout "
                              f"/ its CUDA-event time (average launch of {max(dom['launches'], 1)})",
                "bytes_per_row": dom["bytes_per_row"], "algorithmic_bytes_per_launch": algo_bytes, "launch_ms": launch_ms,
                "share_of_step": dom["share_of_step"],
                "survey_accounting": {"bytes_per_row": bpr_survey,
                                      "achieved": achieved * bpr_survey / dom["bytes_per_row"] if dom_name != "union" else None,
                                      "frac": achieved * bpr_survey / dom["bytes_per_row"] / peak if dom_name != "union" else None,
                                      "note": "SURVEY §8(d) counts the u32 columns (4 + 8F B/row); the device reads fewer bytes for the same rows"},
                "traffic": (tr or {}).get("dram_bytes_per_launch"), "traffic_source": (tr or {}).get("source"),
                "ncu": {k2: tr[k2] for k2 in ("issue_active_pct", "lts_sector_hit_rate_pct", "warp_instructions", "duration_ms_under_ncu",
                                             "captured_at_commit") if tr and k2 in tr} or None,
                "hbm_resident": not image_fits_l2,
                "note": ("the posting image is within ~2x of the 126 MB L2 and the queries are Zipf-drawn: most sectors hit L2, so "
                         "`frac` is layout bytes over the HBM peak, NOT a DRAM utilisation; compare with l2_ceiling_frac and the "
                         "10M-doc configs (cfg3/cfg4), whose image streams from HBM") if image_fits_l2 else
                        "the posting image is far larger than L2: `frac` is an HBM-bandwidth fraction",
                "l2_ceiling_frac": achieved / ceilings["l2_read_gbs"] if ceilings.get("l2_read_gbs") else None,
                "whole_step": {"layout_gbs": sum((c["rows"] or 0) * (c["bytes_per_row"] or 0) for c in classes.values())
                               / (dev_ms / args.steps * 1e-3) / 1e9 if dev_ms else None},
                "classes": classes, "measured_stream_ceilings": ceilings}
    if roofline["whole_step"]["layout_gbs"]:
        roofline["whole_step"]["frac"] = roofline["whole_step"]["layout_gbs"] / peak

    # ---- single-query latency (the reference's interactive use): Q = 1 through pb_query_batch, host to host
    latency = None
    if rank == 0 and world == 1 and not args.no_latency:
        nl = min(200, nq_local)
        ts = []
        for q in range(nl + 5):
            one = fq.slice(q % nl, q % nl + 1)
            t1 = time.perf_counter()
            ix.query_batch_flat(one, calc, cfg.boosts, k)
            ts.append(time.perf_counter() - t1)
        ts = np.asarray(ts[5:]) * 1e6
        latency = {"queries": nl, "p50_us": float(np.percentile(ts, 50)), "p99_us": float(np.percentile(ts, 99)),
                   "mean_us": float(ts.mean()), "launches_per_query": ix.last_stats()["gpu_launches"],
                   "what": "pb_query_batch with ONE query, host buffers in and out (top_k results), wall clock"}
        # the same calls from 4 host threads at once: the library lends each an internal batch (own stream), so they overlap
        import threading
        ones = [fq.slice(q, q + 1) for q in range(nl)]
        for nthr in (1, 4):
            def worker(t):
                for q in range(t, nl, nthr):
                    ix.query_batch_flat(ones[q], calc, cfg.boosts, k)
            for rep in range(2):                       # the first pass creates the borrowed batches and their workspaces
                th = [threading.Thread(target=worker, args=(t,)) for t in range(nthr)]
                t1 = time.perf_counter()
                for x in th:
                    x.start()
                for x in th:
                    x.join()
                latency[f"qps_{nthr}_threads"] = nl / (time.perf_counter() - t1)

    # ---- CPU baseline (rank 0, N = 1 only): the oracle on the host cores, bounded sample
    cpu = None
    want_cpu = (cfg.cfg < 3 or args.cpu_baseline) and not args.no_cpu_baseline
    if rank == 0 and world == 1 and want_cpu:
        try:
            threads = os.cpu_count() or 1
            sample_all = fq.slice(0, min(nq_local, 20_000))
            o, run, n = cpu_arm(cfg, wl, sample_all, scorer_id, threads, target_s=15.0)
            sfq = sample_all.slice(0, n)
            r = run(sfq)
            sb = DeviceBatch(ix, sfq, calc, cfg.boosts, top_k=k)
            sb.run()
            srows = sb.stats()["rows_scored"]
            gres = sb.fetch()
            par = check_against(gres, r, n)
            n1 = max(1, n // 16)
            r1 = run(sfq.slice(0, n1), 1)
            cpu = {"value": srows / r["seconds"], "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"first {n} queries of the same batch ({srows} de-duplicated rows, {r['score_calls']} "
                             f"reference pointer visits), oracle/probly_oracle.cpp on {threads} host threads",
                   "queries_per_sec": n / r["seconds"],
                   "single_thread_pointer_visits_per_sec": r1["score_calls"] / r1["seconds"],
                   "single_thread_mean_query_us": 1e6 * r1["seconds"] / n1,
                   "parity_with_gpu_on_sample": par}
            sb.close()
        except Exception as e:  # the baseline is reported, never the thing measured
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    elif rank == 0 and not want_cpu:
        cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port",
               "sample": "skipped: a 10M-doc oracle index takes longer to build than the bench may run; parity of this line is "
                         "checked against tests/golden/ (see `parity`), the CPU path is timed on cfg1/cfg2 (--cpu-baseline forces it)"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": warmup, "ms_per_step": 1e3 * timed_s / args.steps, "higher_is_better": True,
                "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "queries_per_sec": q_all * args.steps / timed_s,
                "pointer_visits_per_sec": ptr_all * args.steps / timed_s,
                "wall_ms_per_step": 1e3 * wall_max / args.steps, "device_ms_per_step": 1e3 * dev_s / args.steps,
                "timing": "CUDA events on the batch stream around kernels + gather, summed over the steps, max over ranks; "
                          "wall_ms_per_step adds the L2 flush and the host loop",
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": 1e3 * float(ev[0]) / args.steps, "steps": args.steps,
                        "queries_per_sec": q_all * args.steps / float(ev[0]),
                        "what": "pb_batch_reload (pinned host queries -> HBM) + pb_batch_run (kernels + gather) + pb_batch_fetch "
                                "(results -> pinned host), wall clock incl. the L2 flush, max over ranks"},
                "gpu_launches": int(st["gpu_launches"]) * args.steps,
                "stage_ms": {kk: acc[kk] for kk in ("ms_descend", "ms_plan", "ms_score", "ms_side", "ms_finalize", "ms_gather")},
                "per_rank_ms": per_rank_ms,
                "rows": {"scored_per_step": rows_all, "streamed_direct": st["rows_streamed_direct"],
                         "streamed_compact": rows_compact,
                         "streamed_side": st["rows_streamed_side"], "streamed_union": st["rows_streamed_union"],
                         "union_queries": st["union_queries"], "diverted": st["rows_diverted"],
                         "legacy_records": st.get("legacy_records"), "results": res_all, "side_rounds": st["side_rounds"]},
                "parity": parity, "latency_q1": latency, "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    pin.close()
    batch.close()
    del flush


if __name__ == "__main__":
    main()
