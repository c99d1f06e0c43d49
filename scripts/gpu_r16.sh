(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/r16_pytest.log 2>&1
tail -3 gpurun_out/r16_pytest.log
for i in 1 2; do
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r16_bench.json 2> gpurun_out/r16_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r16_bench.json"))
print(round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["stage_ms"].items()}, "e2e", round(d["e2e"]["ms_per_step"],2))
PY
done
