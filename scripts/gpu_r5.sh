(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r5_pytest.log 2>&1
tail -15 gpurun_out/r5_pytest.log
for sh in 0; do
  PB_TAB_REP_SHIFT=$sh timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r5_bench_rep$sh.json 2> gpurun_out/r5_bench_rep$sh.err
  tail -3 gpurun_out/r5_bench_rep$sh.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r5_bench_rep$sh.json"))
print("rep_shift $sh", d["ms_per_step"], d["stage_ms"], d["roofline"]["achieved"], d["e2e"]["ms_per_step"], d["rows"])
PY
done
