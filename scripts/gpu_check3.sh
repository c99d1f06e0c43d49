tag=${1:-chk3}
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/${tag}_pytest.log 2>&1
tail -2 gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-latency 2> gpurun_out/${tag}_bench.err | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); print('cfg1', round(d['ms_per_step'],2), d['parity'])"
grep "resident after" gpurun_out/${tag}_bench.err
