# round-end evidence run on one GPU: tests, bench (with CPU baseline + reference arm), launch list, full capture
tag=${1:-r2}
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/${tag}_pytest.log 2>&1
tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -2 gpurun_out/${tag}_bench.err; cat gpurun_out/${tag}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
cat gpurun_out/${tag}_bench_reference.json
bash scripts/gpu_launchlist.sh ${tag} > /dev/null
bash scripts/gpu_fullcapture.sh ${tag}
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
