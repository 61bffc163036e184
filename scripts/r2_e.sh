set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_pytest_gpu.log; tail -4 gpurun_out/r2e_pytest_gpu.log
timeout 600 python profiles/time_folds.py > gpurun_out/r2e_folds.log 2>&1; grep "default" gpurun_out/r2e_folds.log | grep -v "^{"
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "rc=$?"; tail -5 gpurun_out/r2e_bench.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2e_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')})
    print('roofline',d['roofline']['frac'], d['roofline']['achieved'], 'e2e',d['e2e']['value'])
    print('laplace', d['laplace3d']['value'])
    for k,v in d['extra'].items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ('ms','GB/s','TFLOP/s','Gcells/s','frac_of_measured_hbm_peak','kernel')} if 'check' not in k else '')
except Exception as e: print('parse failed', e)
PY
