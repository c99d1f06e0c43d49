# closing evidence on the final library: cfg2 + cfg4 lines, launch lists of a cfg1 / cfg4-like step
tag=${1:-r3close2}
timeout 900 python bench.py --config cfg2 --steps 3 --warmup 3 > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_cfg1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/${tag}_launches_cfg1_bench.log 2>&1
python profiles/summarize.py launches gpurun_out/${tag}_launches_cfg1.csv "round 3 closing library, cfg1: launch list of warm-up + 1 step + e2e steps" > gpurun_out/${tag}_launches_cfg1.txt 2>&1
head -14 gpurun_out/${tag}_launches_cfg1.txt
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/${tag}_bench_cfg2.json") if l.startswith("{")][-1])
print("cfg2 ms/step", round(d["ms_per_step"],1), "value %.4g"%d["value"], "e2e ms", round(d["e2e"]["ms_per_step"],1), "parity", d["parity"], "lat", d["latency_q1"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"].get("parity_with_gpu_on_sample"))
PY
