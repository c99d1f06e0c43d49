(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_goldens.py -m gpu -x -q 2>&1 | tail -4) > gpurun_out/r11_pytest.log 2>&1
tail -2 gpurun_out/r11_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r11_bench.json 2> gpurun_out/r11_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r11_bench.json"))
print(round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["stage_ms"].items()}, round(d["e2e"]["ms_per_step"],2))
PY
bash scripts/gpu_launchlist.sh r11 
python profiles/summarize.py launches gpurun_out/r11_launches.csv "r11" | head -24
