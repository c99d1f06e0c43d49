# round-3 closing evidence on ONE GPU: everything the final profiles/ files are cut from.
tag=${1:-r3final}
bash scripts/gpu_final.sh $tag
timeout 300 python bench.py --config cfg0,cfg0_bench --steps 20 --warmup 3 > gpurun_out/${tag}_bench_cfg0.json 2> gpurun_out/${tag}_bench_cfg0.err
python - <<PY
import json
for l in open("gpurun_out/${tag}_bench_cfg0.json"):
    if l.startswith("{"):
        d=json.loads(l)
        print(d["config"]["workload"][:40], "ms/step", round(d["ms_per_step"],3), "value %.4g"%d["value"], "e2e ms", round(d["e2e"]["ms_per_step"],3), "parity", d.get("parity"), "lat", (d.get("latency_q1") or {}).get("p50_us"), "cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("parity_with_gpu_on_sample"))
PY
tail -3 gpurun_out/${tag}_bench_cfg0.err
# the default dense-union kernel of a cfg2 step (4000-query batch) under ncu --set full
timeout 900 ncu --set full --clock-control none --import-source on -k regex:union_ -s 2 -c 2 -f -o gpurun_out/${tag}_union_cfg2 python bench.py --config cfg2 --queries 4000 --steps 1 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/${tag}_ncu_cfg2_bench.log 2>&1
# HBM-resident image (cfg3: 10 M docs, 2.1 GB of posting columns): the two scoring launches of one step under ncu --set full
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 6 -c 2 -f -o gpurun_out/${tag}_score_cfg3 python bench.py --config cfg3 --queries 125000 --steps 1 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/${tag}_ncu_cfg3_bench.log 2>&1
ls -la gpurun_out/${tag}_*.ncu-rep
# and the same two 10 M-doc configs un-profiled on this commit (one GPU on its 125 k-query share)
timeout 1500 python bench.py --config cfg3,cfg4 --queries 125000 --steps 5 --warmup 3 --no-latency > gpurun_out/${tag}_bench_cfg34.json 2> gpurun_out/${tag}_bench_cfg34.err
python - <<PY
import json
for l in open("gpurun_out/${tag}_bench_cfg34.json"):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]
        print(d["config"]["workload"][:30], "ms/step", round(d["ms_per_step"],2), "value %.4g"%d["value"], "parity", d.get("parity"), r["class"], round(r["frac"],3), "whole", r.get("whole_step"))
PY
