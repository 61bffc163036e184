set -x
export REPS=2
W="dot sum cols rows gemv32 gevm32 gemv64 gevm64 map mapv stencil"
python profiles/prof_r2.py $W > gpurun_out/r2_prof_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:'sfold|smap|stencil_tma|racu_jit' -o /tmp/r02_kernels_full python profiles/prof_r2.py $W > gpurun_out/r2_prof_ncu.log 2>&1
tail -3 gpurun_out/r2_prof_ncu.log
ncu -i /tmp/r02_kernels_full.ncu-rep --page raw --csv > gpurun_out/r02_kernels_full_raw.csv
ls -la /tmp/r02_kernels_full.ncu-rep
[ $(stat -c %s /tmp/r02_kernels_full.ncu-rep) -lt 50000000 ] && cp /tmp/r02_kernels_full.ncu-rep gpurun_out/
CMD="python bench.py --steps 2 --warmup 1 --extras 0 --laplace 0 --e2e-steps 0"
$CMD > gpurun_out/r2_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_bench.csv $CMD > gpurun_out/r2_bench_ncu.log 2>&1
ls -la gpurun_out/
