# union tests; phase-cycle profile (uprof variant); cfg2 bench line
tag=${1:-r3f}
(timeout 900 python -m pytest tests/test_gpu_union.py -m gpu -x -q 2>&1 | tail -30) > gpurun_out/${tag}_pytest.log 2>&1
tail -3 gpurun_out/${tag}_pytest.log
PB_UNION_PROF=1 PB_LIB_PATH=$PWD/probly_search_b200/_lib/libprobly_b200_uprof.so timeout 600 python bench.py --config cfg2 --queries 20000 --steps 2 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/${tag}_prof.json 2> gpurun_out/${tag}_prof.err
grep "union kernel cycles" gpurun_out/${tag}_prof.err | tail -1
timeout 900 python bench.py --config cfg2 --steps 3 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err
tail -2 gpurun_out/${tag}_bench_cfg2.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_bench_cfg2.json"))
    print("cfg2 ms/step", d["ms_per_step"], "value", d["value"], "stage", d["stage_ms"], "parity", d["parity"])
    print({k:(round(v["ms"],1), v.get("rows_per_sec")) for k,v in d["roofline"]["classes"].items()})
except Exception as e:
    print("failed", e)
PY
