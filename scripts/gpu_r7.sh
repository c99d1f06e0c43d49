PB_DEBUG=1 PB_TAB_REP_SHIFT=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r7_bench.json 2> gpurun_out/r7_bench.err
grep "\[pb\]" gpurun_out/r7_bench.err | tail -2
python - <<PY
import json
d=json.load(open("gpurun_out/r7_bench.json"))
print(d["ms_per_step"], d["stage_ms"], d["roofline"]["achieved"], d["e2e"]["ms_per_step"], d["rows"])
PY
