set -x
timeout 500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['e2e'], d['laplace3d'] and {k:d['laplace3d'][k] for k in ('value','ms_per_step','n_gpus')})
PY
B="python bench.py --steps 2 --warmup 3 --extras 0 --e2e-steps 0 --cpu-log2n 18"
$B > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01c_launches.csv $B > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log
