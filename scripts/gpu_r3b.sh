# round-3 run B (2 GPUs): multi-GPU parity tests, then the weak-scaling bench line at N = 2
tag=${1:-r3b}; n=${2:-2}
(timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/${tag}_pytest_multi.log 2>&1
tail -5 gpurun_out/${tag}_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 8 --warmup 3 > gpurun_out/${tag}_bench_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err
grep "\[bench\]" gpurun_out/${tag}_bench_${n}gpu.err | tail -4; cat gpurun_out/${tag}_bench_${n}gpu.json
