# prefetch / load-policy variants of the scoring kernel (cfg1 bench, no CPU baseline)
for v in base ld1 pf1 pf1ld1 pf1ld1d2 pf2; do
  if [ $v = base ]; then unset PB_LIB_PATH; else export PB_LIB_PATH=$PWD/probly_search_b200/_lib/libprobly_b200_$v.so; fi
  PB_TAB_REP_SHIFT=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r4_$v.json 2> gpurun_out/r4_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r4_$v.json"))
    print("$v", round(d["ms_per_step"],2), {k:round(x,2) for k,x in d["stage_ms"].items()}, round(d["roofline"]["achieved"]), round(d["e2e"]["ms_per_step"],2))
except Exception as e:
    print("$v failed", e)
PY
done
