# round-3 run E (1 GPU): union tests; phase-cycle profile of the union kernel; ncu full capture of it on a short batch
tag=${1:-r3e}
(timeout 900 python -m pytest tests/test_gpu_union.py -m gpu -x -q 2>&1 | tail -30) > gpurun_out/${tag}_pytest.log 2>&1
tail -3 gpurun_out/${tag}_pytest.log
PB_UNION_PROF=1 PB_LIB_PATH=$PWD/probly_search_b200/_lib/libprobly_b200_uprof.so timeout 600 python bench.py --config cfg2 --queries 20000 --steps 2 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/${tag}_prof.json 2> gpurun_out/${tag}_prof.err
grep "union kernel cycles" gpurun_out/${tag}_prof.err | tail -2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:union_kernel -s 2 -c 1 -o gpurun_out/${tag}_union python bench.py --config cfg2 --queries 4000 --steps 1 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/${tag}_ncu_bench.log 2>&1
ls -la gpurun_out/${tag}_union.ncu-rep
