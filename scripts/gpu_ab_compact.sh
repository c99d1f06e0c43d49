# A/B of the compact tile copy (PB_POSTING_COMPACT): cfg3 + cfg4 on one GPU (125 k-query share), cfg1 forced on
for v in 0 1; do
  PB_POSTING_COMPACT=$v timeout 1200 python bench.py --config cfg3,cfg4 --queries 125000 --steps 5 --warmup 3 --no-latency > gpurun_out/abc_cfg34_$v.json 2> gpurun_out/abc_cfg34_$v.err
  PB_POSTING_COMPACT=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-latency --no-cpu-baseline > gpurun_out/abc_cfg1_$v.json 2> gpurun_out/abc_cfg1_$v.err
done
python - <<PY
import json
for f in ("cfg34_0","cfg34_1","cfg1_0","cfg1_1"):
    try:
        for l in open("gpurun_out/abc_%s.json"%f):
            if not l.startswith("{"): continue
            d=json.loads(l); r=d["roofline"]
            print(f, d["config"]["workload"][:5], "ms/step", round(d["ms_per_step"],2), "parity", d.get("parity"), {k:round(v["ms"],2) for k,v in r["classes"].items() if v["ms"]>0.01},
                  "compact rows", d["rows"].get("streamed_compact"), "direct B/row", round(r["classes"]["direct"]["bytes_per_row"],3), "direct GB/s", round(r["classes"]["direct"].get("layout_gbs",0)))
    except Exception as e:
        print(f, "failed", e)
PY
tail -2 gpurun_out/abc_cfg34_1.err
