# round-3 profile + sanitizer evidence (1 GPU).  Numbers printed under a profiler are never bench values.
tag=${1:-r3p}
# 1. launch lists (cold-cache, serialised: compare SHARES)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_cfg1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/${tag}_launches_cfg1_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_cfg2.csv python bench.py --config cfg2 --queries 20000 --steps 1 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/${tag}_launches_cfg2_bench.log 2>&1
# 2. full captures: the two scoring launches of a cfg1 step (class S, class G); the union kernels of a cfg2 step
timeout 900 ncu --set full --import-source on --clock-control none -k regex:score_kernel -s 6 -c 2 -f -o gpurun_out/${tag}_score_cfg1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/${tag}_ncu_cfg1_bench.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:union_kernel -s 6 -c 2 -f -o gpurun_out/${tag}_union_cfg2 python bench.py --config cfg2 --queries 4000 --steps 1 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/${tag}_ncu_cfg2_bench.log 2>&1
ls -la gpurun_out/${tag}_*.ncu-rep
# 3. sanitizers on the new code (union kernel: shared-memory atomics / lists; gather path)
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest "tests/test_gpu_union.py::test_term_pools_and_consumed_query_terms" "tests/test_gpu_union.py::test_queries_outside_the_dense_envelope_take_the_list_route" "tests/test_gpu_union.py::test_random_corpora_through_both_routes" "tests/test_gpu_parity.py::test_docs_with_many_events" "tests/test_gpu_multi.py::test_world1_gather_matches_oracle_and_local_fetch" -m gpu -x -q > gpurun_out/${tag}_memcheck.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${tag}_memcheck.log | head -5
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 20 python -m pytest "tests/test_gpu_union.py::test_term_pools_and_consumed_query_terms" "tests/test_gpu_union.py::test_random_corpora_through_both_routes[0-union_all]" "tests/test_gpu_union.py::test_random_corpora_through_both_routes[3-union_all]" -m gpu -x -q > gpurun_out/${tag}_racecheck.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" gpurun_out/${tag}_racecheck.log | head -5
timeout 420 compute-sanitizer --tool initcheck --error-exitcode 99 --print-limit 20 python -m pytest "tests/test_gpu_union.py::test_term_pools_and_consumed_query_terms" "tests/test_gpu_goldens.py" -m gpu -x -q -k "pools or bm25_one_field or zero_to_one_case0" > gpurun_out/${tag}_initcheck.log 2>&1
echo "initcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Uninitialized" gpurun_out/${tag}_initcheck.log | head -5
