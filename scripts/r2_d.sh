set -x
timeout 1500 python -m pytest tests/test_gpu_configs.py -m gpu -x -q > gpurun_out/r2d_pytest_cfg.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_pytest_cfg.log; tail -25 gpurun_out/r2d_pytest_cfg.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "rc=$?"; tail -5 gpurun_out/r2d_bench.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2d_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','check')})
    print('roofline',d['roofline']['frac'], 'e2e',d['e2e']['value'])
    print('laplace', d['laplace3d']['value'], d['laplace3d'].get('check'))
    for k,v in d['extra'].items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ('ms','GB/s','TFLOP/s','Gcells/s','frac_of_measured_hbm_peak','kernel','cublas_same_box_TFLOP/s','max_err_over_sum_abs_products')} if 'check' not in k else v)
    print('cpu', json.dumps(d['cpu_baseline'])[:1500])
    print('cxx', json.dumps(d.get('e2e_cxx'))[:800])
except Exception as e: print('parse failed', e)
PY
