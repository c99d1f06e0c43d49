PB_TAB_REP_SHIFT=0 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,smsp__warps_active.avg.per_cycle_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_drain_per_issue_active.ratio --clock-control none -c 60 --csv --log-file gpurun_out/r6_ncu.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r6_ncu_bench.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r6_ncu.csv")) if len(r)>14 and r[0].isdigit()]
from collections import OrderedDict
k=OrderedDict()
for r in rows:
    k.setdefault((int(r[0]), r[4][:60]), {})[r[12]]=r[14]
for (i,name),m in k.items():
    print(i, name, {a.split('.')[0][-28:]:b for a,b in m.items()})
PY
