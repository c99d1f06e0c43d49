# 10M-doc configs on ONE GPU (the per-rank share of the 8-GPU batch): validates the flow + the HBM-resident numbers
tag=${1:-r3n}
timeout 1500 python bench.py --config cfg3,cfg4 --queries 125000 --steps 3 --warmup 3 > gpurun_out/${tag}_bench_cfg34_1gpu.json 2> gpurun_out/${tag}_bench_cfg34_1gpu.err
grep "\[bench\]" gpurun_out/${tag}_bench_cfg34_1gpu.err | tail -5
python - <<PY
import json
for line in open("gpurun_out/${tag}_bench_cfg34_1gpu.json"):
    try:
        d=json.loads(line)
        print(d["config"]["workload"][:40], "ms/step", round(d["ms_per_step"],1), "value %.3g" % d["value"], "stage", {k:round(v,1) for k,v in d["stage_ms"].items()}, "parity", d["parity"])
        r=d["roofline"]; print("  roofline", r["class"], round(r["achieved"]), "GB/s frac", round(r["frac"],3), "whole", r["whole_step"])
    except Exception as e:
        print("failed", e)
PY
