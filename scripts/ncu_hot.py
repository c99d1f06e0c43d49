"""Hot SASS runs of every kernel in an `ncu --page source --csv` export: contiguous instructions with
the same execution count, their share of executed instructions and of stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
detail = len(sys.argv) > 2 and sys.argv[2]
kern, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "ins": []}; kern.append(cur); continue
    if r and r[0] == "Address": cur["hdr"] = r; continue
    if cur is not None and r and r[0].startswith("0x"): cur["ins"].append(r)
seen = set()
for k in kern:
    if k["name"] in seen: continue
    seen.add(k["name"])
    h = k["hdr"]; iS = h.index("Source"); iE = h.index("Instructions Executed"); iW = h.index("Warp Stall Sampling (All Samples)")
    tot = sum(int(r[iE]) for r in k["ins"]); tw = sum(int(r[iW]) for r in k["ins"])
    print("==", k["name"][:70], "instr", tot, "samples", tw, "n_sass", len(k["ins"]))
    runs = []
    for i, r in enumerate(k["ins"]):
        e = int(r[iE]); w = int(r[iW])
        if runs and abs(e - runs[-1]["e"]) <= 0.02 * max(e, runs[-1]["e"], 1): runs[-1]["n"] += 1; runs[-1]["sum"] += e; runs[-1]["w"] += w; runs[-1]["end"] = i
        else: runs.append({"e": e, "n": 1, "sum": e, "w": w, "start": i, "end": i})
    for b in sorted(sorted(runs, key=lambda x: -x["sum"])[:12], key=lambda x: x["start"]):
        print("  sass[%d..%d] n=%d exec=%.3e total=%.3e (%.1f%%) stalls %.1f%%" % (b["start"], b["end"], b["n"], b["e"], b["sum"], 100 * b["sum"] / tot, 100 * b["w"] / tw))
    if detail and detail in k["name"]:
        lo, hi = int(sys.argv[3]), int(sys.argv[4])
        for i in range(lo, hi):
            r = k["ins"][i]
            print(i, r[iS].strip()[:70].ljust(70), r[iE], "%.2f%%" % (100 * int(r[iW]) / tw))
