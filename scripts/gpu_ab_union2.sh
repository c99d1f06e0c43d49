# A/B of the two union kernels / shard widths on ONE box: kernel:wbits pairs
(timeout 900 python -m pytest tests/test_gpu_union.py -m gpu -x -q 2>&1 | tail -8) > gpurun_out/abw_pytest.log 2>&1; tail -3 gpurun_out/abw_pytest.log
for round in 1 2; do
for spec in "$@"; do
  k=${spec%%:*}; w=${spec##*:}
  PB_UNION_KERNEL=$k PB_UNION_WBITS=$w timeout 300 python bench.py --config cfg2 --queries 20000 --steps 3 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/abw_${k}_$w.json 2> gpurun_out/abw_${k}_$w.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/abw_${k}_$w.json") if l.startswith("{")][-1])
    print("$k wbits $w", "step", round(d["ms_per_step"],1), "union", round(d["roofline"]["classes"]["union"]["ms"],1), "parity", d["parity"].get("golden_ok"))
except Exception as e:
    print("$k $w failed", e); print(open("gpurun_out/abw_${k}_$w.err").read()[-600:])
PY
done; done
