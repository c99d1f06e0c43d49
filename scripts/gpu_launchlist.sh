# launch list of one bench step (cold-cache, serialised: compare SHARES) -> gpurun_out/$1_launches.csv
tag=${1:-rX}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_launches_bench.log 2>&1
python profiles/summarize.py gpurun_out/${tag}_launches.csv 2>/dev/null | head -40
