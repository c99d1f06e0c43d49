"""CPU-only: `bench.py --impl reference` (the oracle timed on the host cores) runs without a GPU and
prints the contract's JSON line; the workload generator gives every document its own token stream."""
import json
import os
import subprocess
import sys

import numpy as np

from probly_search_b200 import Index
from probly_search_b200 import workload as W
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_runs_on_cpu():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "cfg0",
                                   "--docs", "4000", "--vocab", "2048", "--queries", "64", "--steps", "1", "--warmup", "0"],
                                  cwd=ROOT, timeout=300)
    line = json.loads(out.decode().strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "scored-postings/sec" and line["unit"] == "postings/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]


def test_documents_do_not_share_a_sliding_token_window():
    """Regression: per-document streams seeded `seed + gamma * doc` made doc d+1's stream doc d's shifted by one
    draw, so a rare term sat in ~28 CONSECUTIVE documents.  With independent streams the docs of a rare
    term are spread: almost no two of them are neighbours."""
    cfg = W.CONFIGS["cfg1"]
    wl = W.Workload(cfg, n_docs=20_000, vocab=1 << 12)
    ix = Index(cfg.n_fields)
    wl.build_into(ix)
    a = H.image_arrays(ix.flatten())
    trb = a["term_row_begin"].astype(np.int64)
    rows = np.diff(trb)
    rare = np.nonzero((rows >= 8) & (rows <= 200))[0]
    assert len(rare) > 200
    adjacent = total = 0
    for t in rare[:400]:
        d = a["post_doc"][trb[t]:trb[t + 1]].astype(np.int64)
        adjacent += int(np.count_nonzero(np.diff(d) == 1))
        total += len(d) - 1
    assert adjacent / total < 0.05, (adjacent, total)
