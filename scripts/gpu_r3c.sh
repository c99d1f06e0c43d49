# round-3 run C (1 GPU): union-path tests, then bench cfg2
tag=${1:-r3c}
(timeout 1200 python -m pytest tests/test_gpu_union.py tests/test_gpu_parity.py tests/test_gpu_goldens.py -m gpu -x -q 2>&1 | tail -30) > gpurun_out/${tag}_pytest.log 2>&1
tail -12 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py --config cfg2 --steps 3 --warmup 3 > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err
tail -3 gpurun_out/${tag}_bench_cfg2.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_bench_cfg2.json"))
    print("cfg2 ms/step", d["ms_per_step"], "value", d["value"], "stage", d["stage_ms"], "rows", d["rows"], "parity", d["parity"])
    print({k:(v["ms"], v.get("rows_per_sec")) for k,v in d["roofline"]["classes"].items()})
    print("cpu", d["cpu_baseline"])
except Exception as e:
    print("failed", e)
PY
