# full GPU suite + smoke + default bench line on the current build
tag=${1:-chk}
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/${tag}_pytest.log 2>&1
tail -3 gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_cfg1.json 2> gpurun_out/${tag}_bench_cfg1.err
timeout 300 python bench.py --config cfg0 --steps 20 --warmup 3 > gpurun_out/${tag}_bench_cfg0.json 2> gpurun_out/${tag}_bench_cfg0.err
python - <<PY
import json
for f in ("cfg1","cfg0"):
    for l in open("gpurun_out/${tag}_bench_%s.json"%f):
        if not l.startswith("{"): continue
        d=json.loads(l)
        print(f, "ms/step", round(d["ms_per_step"],3), "value %.4g"%d["value"], "e2e", round(d["e2e"]["ms_per_step"],3), "parity", d.get("parity"), "lat", d.get("latency_q1"))
PY
