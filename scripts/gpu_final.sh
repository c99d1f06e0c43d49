# round-end evidence on one GPU: full GPU suite, smoke, the default bench line (+ reference arm), cfg2 line
tag=${1:-r3}
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/${tag}_pytest.log 2>&1
tail -3 gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -3 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_cfg1.json 2> gpurun_out/${tag}_bench_cfg1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_cfg1_reference.json 2> gpurun_out/${tag}_bench_cfg1_reference.err
timeout 900 python bench.py --config cfg2 --steps 3 --warmup 3 > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err
python - <<PY
import json
for f in ("cfg1","cfg1_reference","cfg2"):
    try:
        d=json.loads([l for l in open("gpurun_out/${tag}_bench_%s.json"%f) if l.startswith("{")][-1])
        print(f, "ms/step", round(d["ms_per_step"],2), "value %.4g"%d["value"], "e2e", d["e2e"].get("ms_per_step"), "parity", d.get("parity"), "lat", (d.get("latency_q1") or {}).get("p50_us"))
        if "roofline" in d: print("   roofline", d["roofline"]["class"], round(d["roofline"]["frac"],3), {k:round(v["ms"],2) for k,v in d["roofline"]["classes"].items()}, "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
