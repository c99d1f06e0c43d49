# A/B on a DRAM-resident image: cfg4 shape at 4M docs (0.85 GB narrow image, boosts [2, .5], 5 % removed pre-vacuum)
for round in 1 2; do
for v in "$@"; do
  if [ $v = base ]; then unset PB_LIB_PATH; else export PB_LIB_PATH=$PWD/probly_search_b200/_lib/libprobly_b200_$v.so; fi
  timeout 600 python bench.py --config cfg4 --docs 4000000 --queries 50000 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ab4_$v.json 2> gpurun_out/ab4_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab4_$v.json"))
    print("$v", round(d["ms_per_step"],2), {k:round(x,2) for k,x in d["stage_ms"].items()}, "rows/s S", round(d["roofline"]["rows_per_sec_this_launch"]/1e9,1), "G")
except Exception as e:
    print("$v failed", e)
PY
done; done
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_goldens.py -m gpu -x -q 2>&1 | tail -3)
unset PB_LIB_PATH
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cfg1', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms'].items()}, 'e2e', round(d['e2e']['ms_per_step'],2))"
