tag=${1:-r3g}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:union_ -s 2 -c 1 -o gpurun_out/${tag}_union python bench.py --config cfg2 --queries 4000 --steps 1 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/${tag}_ncu_bench.log 2>&1
ls -la gpurun_out/${tag}_union.ncu-rep
