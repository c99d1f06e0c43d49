(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r8_pytest.log 2>&1
tail -3 gpurun_out/r8_pytest.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:score_kernel -c 2 -f -o gpurun_out/r8_score python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r8_ncu_bench.log 2>&1
ls -la gpurun_out/
