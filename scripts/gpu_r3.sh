set -x
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r3_pytest.log 2>&1
tail -3 gpurun_out/r3_pytest.log
for sh in 0 3 4; do
  PB_TAB_REP_SHIFT=$sh timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r3_bench_rep$sh.json 2> gpurun_out/r3_bench_rep$sh.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r3_bench_rep$sh.json"))
print("rep_shift $sh", d["ms_per_step"], d["stage_ms"], d["roofline"]["achieved"], d["e2e"]["ms_per_step"])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum --clock-control none -k regex:score_kernel -c 4 --csv --log-file gpurun_out/r3_ncu_metrics.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r3_ncu_bench.log 2>&1
tail -30 gpurun_out/r3_ncu_metrics.csv | cut -c1-300
