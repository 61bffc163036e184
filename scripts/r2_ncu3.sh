set -x
CMD="python bench.py --steps 2 --warmup 1 --extras 0 --laplace 0 --e2e-steps 0"
$CMD > gpurun_out/r2d_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02d_ncu_launches_bench.csv $CMD > gpurun_out/r2d_bench_ncu.log 2>&1
tail -2 gpurun_out/r2d_bench_ncu.log | cut -c1-200
