tag=${1:-chk2}
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/${tag}_pytest.log 2>&1
tail -2 gpurun_out/${tag}_pytest.log
python scripts/q1_probe.py cfg0 2>&1 | tail -1
python scripts/q1_probe.py cfg1 200000 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 600 --launch-count 42 --csv --log-file gpurun_out/${tag}_q1_launches.csv python scripts/q1_probe.py cfg0 > /dev/null 2>&1
python profiles/summarize.py launches gpurun_out/${tag}_q1_launches.csv "Q=1 on cfg0: 42 consecutive launches (7 queries)" 2>&1 | head -20
