# 8 GPUs: multi-GPU parity tests, cfg1 weak-scaling line, then the two 8-GPU configurations (cfg3, cfg4: 10M docs, 1M queries sharded)
tag=${1:-r3o}; n=${2:-8}
free -g | head -2 > gpurun_out/${tag}_host.txt; nproc >> gpurun_out/${tag}_host.txt
(timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/${tag}_pytest_multi.log 2>&1
tail -3 gpurun_out/${tag}_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 8 --warmup 3 > gpurun_out/${tag}_bench_cfg1_${n}gpu.json 2> gpurun_out/${tag}_bench_cfg1_${n}gpu.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_bench_cfg1_${n}gpu.json"))
    print("cfg1 x$n ms/step", round(d["ms_per_step"],2), "value %.4g" % d["value"], "e2e ms", round(d["e2e"]["ms_per_step"],2), "parity", d["parity"], "per-rank", d["per_rank_ms"])
except Exception as e:
    print("cfg1 failed", e)
PY
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --config cfg3,cfg4 --steps 5 --warmup 3 > gpurun_out/${tag}_bench_cfg34_${n}gpu.json 2> gpurun_out/${tag}_bench_cfg34_${n}gpu.err
grep "\[bench\] rank 0" gpurun_out/${tag}_bench_cfg34_${n}gpu.err | tail -4
python - <<PY
import json
for line in open("gpurun_out/${tag}_bench_cfg34_${n}gpu.json"):
    try:
        d=json.loads(line)
        print(d["config"]["workload"][:30], "x", d["n_gpus"], "ms/step", round(d["ms_per_step"],1), "value %.4g" % d["value"], "q/s %.4g" % d["queries_per_sec"], "e2e ms", round(d["e2e"]["ms_per_step"],1), "parity", d["parity"])
        print("   per-rank [total, score, side, gather]", d["per_rank_ms"])
    except Exception as e:
        print("failed", e)
PY
