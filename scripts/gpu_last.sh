(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/r2b_pytest.log 2>&1
tail -2 gpurun_out/r2b_pytest.log
timeout 300 python bench.py --steps 8 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
cat gpurun_out/r2b_bench.json | cut -c1-400
