for v in base t1024; do
  if [ $v = base ]; then unset PB_LIB_PATH; else export PB_LIB_PATH=$PWD/probly_search_b200/_lib/libprobly_b200_$v.so; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r15_$v.json 2> gpurun_out/r15_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r15_$v.json"))
    print("$v", round(d["ms_per_step"],2), {k:round(x,2) for k,x in d["stage_ms"].items()}, round(d["e2e"]["ms_per_step"],2), d["roofline"]["measured_stream_ceilings"])
except Exception as e:
    print("$v failed", e)
PY
done
