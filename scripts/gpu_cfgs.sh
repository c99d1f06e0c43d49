# other BASELINE configs on one GPU (no CPU baseline: the oracle index of a 10M-doc corpus takes minutes to build)
timeout 900 python bench.py --config cfg2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_cfg2.json 2> gpurun_out/r2a_bench_cfg2.err
tail -2 gpurun_out/r2a_bench_cfg2.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2a_bench_cfg2.json"))
print("cfg2", round(d["ms_per_step"],2), d["value"], d["queries_per_sec"], {k:round(v,2) for k,v in d["stage_ms"].items()}, d["rows"])
PY
timeout 1500 python bench.py --config cfg4 --queries 100000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_cfg4.json 2> gpurun_out/r2a_bench_cfg4.err
tail -2 gpurun_out/r2a_bench_cfg4.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2a_bench_cfg4.json"))
print("cfg4", round(d["ms_per_step"],2), d["value"], d["queries_per_sec"], {k:round(v,2) for k,v in d["stage_ms"].items()}, d["rows"], d["roofline"]["achieved"])
PY
