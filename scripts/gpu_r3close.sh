# last 1-GPU check of the round: smoke, default bench line + reference arm, cfg0 / cfg0_bench lines
tag=${1:-r3close}
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -3 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_cfg1.json 2> gpurun_out/${tag}_bench_cfg1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_cfg1_reference.json 2> gpurun_out/${tag}_bench_cfg1_reference.err
timeout 300 python bench.py --config cfg0,cfg0_bench --steps 20 --warmup 3 > gpurun_out/${tag}_bench_cfg0.json 2> gpurun_out/${tag}_bench_cfg0.err
python - <<PY
import json
for f in ("cfg1","cfg1_reference","cfg0"):
    for l in open("gpurun_out/${tag}_bench_%s.json"%f):
        if not l.startswith("{"): continue
        d=json.loads(l)
        print(f, d["config"]["workload"][:12], "ms/step", round(d["ms_per_step"],3), "value %.4g"%d["value"], "e2e", d["e2e"].get("ms_per_step"), "parity", d.get("parity"), "lat", (d.get("latency_q1") or {}).get("p50_us"), "cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("parity_with_gpu_on_sample"))
PY
