(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/r14_pytest.log 2>&1
tail -2 gpurun_out/r14_pytest.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r14_bench.json 2> gpurun_out/r14_bench.err
tail -2 gpurun_out/r14_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r14_bench.json"))
print(round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["stage_ms"].items()}, "e2e", round(d["e2e"]["ms_per_step"],2))
print(json.dumps(d["roofline"]))
print(json.dumps(d["cpu_baseline"]))
PY
