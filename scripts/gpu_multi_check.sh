# multi-GPU regression on the current build: the -m gpu multi tests, then the default bench under torchrun
n=${1:-2}; tag=${2:-mchk}
(timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -6) > gpurun_out/${tag}_pytest_multi.log 2>&1
tail -3 gpurun_out/${tag}_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${tag}_bench_cfg1_${n}gpu.json 2> gpurun_out/${tag}_bench_cfg1_${n}gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus $n --steps 1 --warmup 1 > gpurun_out/${tag}_bench_ref_${n}gpu.json 2> gpurun_out/${tag}_bench_ref_${n}gpu.err
python - <<PY
import json
for f in ("cfg1","ref"):
    try:
        d=json.loads([l for l in open("gpurun_out/${tag}_bench_%s_${n}gpu.json"%f) if l.startswith("{")][-1])
        print(f, "x$n ms/step", round(d["ms_per_step"],2), "value %.4g" % d["value"], "e2e", d["e2e"], "parity", d.get("parity"), "per-rank", d.get("per_rank_ms"))
    except Exception as e:
        print(f, "failed", e)
PY
