# compute-sanitizer on the code added late in round 3: row bits for removed docs, compact tiles, boost fold, borrowed batches
tag=${1:-san3}
T='tests/test_gpu_parity.py::test_removed_docs_by_row_bits_and_by_bitmap_probe[cfg4-cfg4-1] tests/test_gpu_parity.py::test_removed_docs_by_row_bits_and_by_bitmap_probe[cfg2-cfg4-1] tests/test_gpu_parity.py::test_power_of_two_boosts_folded_into_the_table[boosts2] tests/test_gpu_compact.py::test_compact_tiles_with_non_unit_boosts_and_full_results tests/test_gpu_parity.py::test_one_call_queries_from_several_host_threads'
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest $T -m gpu -x -q > gpurun_out/${tag}_memcheck.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${tag}_memcheck.log | head -5
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 20 python -m pytest "tests/test_gpu_parity.py::test_removed_docs_by_row_bits_and_by_bitmap_probe[cfg4-cfg4-1]" tests/test_gpu_compact.py::test_compact_tiles_with_non_unit_boosts_and_full_results -m gpu -x -q > gpurun_out/${tag}_racecheck.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" gpurun_out/${tag}_racecheck.log | head -5
