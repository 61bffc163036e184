set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2r_pytest_gpu.log; tail -4 gpurun_out/r2r_pytest_gpu.log
timeout 1200 python bench.py > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "rc=$?"; tail -5 gpurun_out/r2r_bench.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2r_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')})
    print('roofline',d['roofline']['frac'], d['roofline']['achieved'], 'e2e',d['e2e']['value'])
    print('laplace', d['laplace3d']['value'], d['laplace3d'].get('two_steps_per_pass'))
    for k,v in d['extra'].items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ('ms','GB/s','TFLOP/s','Gcells/s','frac_of_measured_hbm_peak','kernel','cublas_TFLOP/s')} if 'check' not in k else '')
except Exception as e: print('parse failed', e)
PY
