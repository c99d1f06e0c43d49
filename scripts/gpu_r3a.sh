# round-3 run A (1 GPU): GPU tests, then the default bench line
tag=${1:-r3a}
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/${tag}_pytest.log 2>&1
tail -5 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -3 gpurun_out/${tag}_bench.err; cat gpurun_out/${tag}_bench.json
