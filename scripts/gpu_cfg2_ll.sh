timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/cfg2_launches.csv python bench.py --config cfg2 --queries 10000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/cfg2_launches_bench.log 2>&1
python profiles/summarize.py launches gpurun_out/cfg2_launches.csv "cfg2 10k queries" | head -16
timeout 300 python bench.py --config cfg2 --queries 10000 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],1), d['stage_ms'], d['rows'])"
