# A/B of library variants on ONE box: bash scripts/gpu_ab.sh base t768 ...   (alternating, 2 rounds)
for round in 1 2; do
for v in "$@"; do
  if [ $v = base ]; then unset PB_LIB_PATH; else export PB_LIB_PATH=$PWD/probly_search_b200/_lib/libprobly_b200_$v.so; fi
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/ab_$v.json") if l.startswith('{')][-1])
    print("$v", round(d["ms_per_step"],2), {k:round(x,2) for k,x in d["stage_ms"].items()}, "e2e", round(d["e2e"]["ms_per_step"],2), {k:round(v["ms"],2) for k,v in d["roofline"]["classes"].items() if v["ms"]>0.01}, "parity", d["parity"].get("golden_ok"))
except Exception as e:
    print("$v failed", e)
PY
done; done
