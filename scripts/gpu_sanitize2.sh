# WARNING: the initcheck leg below took ~25 GPU-MINUTES for 30 small tests on the round-1 box (memcheck and
# racecheck take seconds).  Run it on a handful of tests only, with a short --timeout on gpurun.
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest "tests/test_gpu_parity.py::test_scaled_configs_both_layouts" "tests/test_gpu_parity.py::test_index_served_from_an_image_file" -m gpu -x -q > gpurun_out/sanitize_memcheck_scaled.log 2>&1
echo "memcheck scaled exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_memcheck_scaled.log | head
timeout 1500 compute-sanitizer --tool initcheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_goldens.py "tests/test_gpu_parity.py::test_random_small_corpora" "tests/test_gpu_parity.py::test_rank_directory_for_every_list" -m gpu -x -q > gpurun_out/sanitize_initcheck.log 2>&1
echo "initcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Uninitialized" gpurun_out/sanitize_initcheck.log | head
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_goldens.py -m gpu -x -q > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" gpurun_out/sanitize_racecheck.log | head
