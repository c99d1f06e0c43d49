# one ncu --set full capture of the two scoring launches (class S, class G) of a cfg1 batch
tag=${1:-rX}
timeout 900 ncu --set full --import-source on --clock-control none -k regex:score_kernel -c 2 -f -o gpurun_out/${tag}_score python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
ls -la gpurun_out/${tag}_score.ncu-rep
