# A/B of union-kernel shapes on ONE box: variant:wbits pairs
for round in 1 2; do
for spec in "$@"; do
  v=${spec%%:*}; w=${spec##*:}
  if [ $v = base ]; then unset PB_LIB_PATH; else export PB_LIB_PATH=$PWD/probly_search_b200/_lib/libprobly_b200_$v.so; fi
  PB_UNION_WBITS=$w timeout 300 python bench.py --config cfg2 --queries 20000 --steps 3 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/abu_${v}_$w.json 2> gpurun_out/abu_${v}_$w.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/abu_${v}_$w.json"))
    print("$v wbits $w", "step", round(d["ms_per_step"],1), "union", round(d["roofline"]["classes"]["union"]["ms"],1), "parity", d["parity"].get("golden_ok"))
except Exception as e:
    print("$v $w failed", e)
PY
done; done
