tag=${1:-r3q}
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/${tag}_pytest.log 2>&1
tail -6 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${tag}_bench.json") if l.startswith("{")][-1])
    print("cfg1 ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), "parity", d["parity"], "latency", d["latency_q1"])
except Exception as e:
    print("cfg1 failed", e)
PY
for div in 16 64 0; do
  PB_UNION_MIN_DIV=$div timeout 300 python bench.py --config cfg2 --queries 20000 --steps 3 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/${tag}_cfg2_div$div.json 2> gpurun_out/${tag}_cfg2_div$div.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${tag}_cfg2_div$div.json") if l.startswith("{")][-1])
    c=d["roofline"]["classes"]
    print("cfg2 20k div $div: step", round(d["ms_per_step"],1), "union", round(c["union"]["ms"],1), "side", round(c["side_score"]["ms"]+c["side_mark"]["ms"]+c["side_fold"]["ms"],1), "uq", d["rows"]["union_queries"], "parity", d["parity"].get("golden_ok"))
except Exception as e:
    print("cfg2 div $div failed", e)
PY
done
