(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r9_pytest_narrow.log 2>&1
tail -3 gpurun_out/r9_pytest_narrow.log
(PB_POSTING_LAYOUT=wide timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_goldens.py -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r9_pytest_wide.log 2>&1
tail -3 gpurun_out/r9_pytest_wide.log
for lay in auto wide; do
for sh in 0 4; do
  PB_POSTING_LAYOUT=$lay PB_TAB_REP_SHIFT=$sh timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r9_bench_${lay}_rep$sh.json 2> gpurun_out/r9_bench_${lay}_rep$sh.err
  python - <<PY
import json
try:
  d=json.load(open("gpurun_out/r9_bench_${lay}_rep$sh.json"))
  print("$lay rep_shift $sh", round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["stage_ms"].items()}, round(d["e2e"]["ms_per_step"],2))
except Exception as e: print("$lay $sh failed", e)
PY
done; done
