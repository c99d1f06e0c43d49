# full GPU suite on the new build, then A/B of the power-of-two boost fold (PB_FOLD_BOOST) on cfg4 (one GPU, 125 k-query share)
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/abf_pytest.log 2>&1
tail -3 gpurun_out/abf_pytest.log
for v in 0 1; do
  PB_FOLD_BOOST=$v timeout 1200 python bench.py --config cfg3,cfg4 --queries 125000 --steps 5 --warmup 3 --no-latency > gpurun_out/abf_cfg34_$v.json 2> gpurun_out/abf_cfg34_$v.err
done
python - <<PY
import json
for f in ("cfg34_0","cfg34_1"):
    try:
        for l in open("gpurun_out/abf_%s.json"%f):
            if not l.startswith("{"): continue
            d=json.loads(l); r=d["roofline"]
            print(f, d["config"]["workload"][:5], "ms/step", round(d["ms_per_step"],2), "parity", d.get("parity"), {k:round(v["ms"],2) for k,v in r["classes"].items() if v["ms"]>0.01}, "frac", round(r["frac"],3), "whole", round(r["whole_step"]["frac"],3))
    except Exception as e:
        print(f, "failed", e)
PY
