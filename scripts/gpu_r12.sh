M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__occupancy_limit_registers,smsp__cycles_active.avg,sm__cycles_active.avg,sm__cycles_active.max,sm__cycles_active.min,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,smsp__thread_inst_executed_per_inst_executed.ratio
for s in long_scoreboard short_scoreboard wait lg_throttle no_instruction branch_resolving membar; do M=$M,smsp__average_warps_issue_stalled_${s}_per_issue_active.ratio; done
timeout 600 ncu --metrics $M --clock-control none -k regex:mark_kernel -c 2 --csv --log-file gpurun_out/r12_ncu.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r12_ncu_bench.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r12_ncu.csv")) if len(r)>14 and r[0].isdigit()]
k={}
for r in rows: k.setdefault((int(r[0]), r[4][:40]), {})[r[12]]=r[14]
for (i,name),m in sorted(k.items()):
    print(i,name)
    for a,b in m.items():
        a=a.replace("smsp__average_warps_issue_stalled_","stall:").replace("_per_issue_active.ratio","")
        print("    ",a,b)
PY
