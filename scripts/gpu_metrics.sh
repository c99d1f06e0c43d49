# ncu metric subset (time, instructions, issue, stall reasons, shared-memory conflicts) of the two scoring launches
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
for s in long_scoreboard short_scoreboard wait math_pipe_throttle not_selected branch_resolving dispatch_stall mio_throttle lg_throttle no_instruction; do M=$M,smsp__average_warps_issue_stalled_${s}_per_issue_active.ratio; done
timeout 600 ncu --metrics $M --clock-control none -k regex:${1:-score_kernel} -c 2 --csv --log-file gpurun_out/metrics_ncu.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/metrics_ncu_bench.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/metrics_ncu.csv")) if len(r)>14 and r[0].isdigit()]
k={}
for r in rows: k.setdefault((int(r[0]), r[4][:40]), {})[r[12]]=r[14]
for (i,name),m in sorted(k.items()):
    print(i,name)
    for a,b in m.items():
        print("    ",a.replace("smsp__average_warps_issue_stalled_","stall:").replace("_per_issue_active.ratio",""),b)
PY
