set -x
timeout 1800 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/r2z_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2z_pytest_gpu.log; tail -20 gpurun_out/r2z_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; tail -2 gpurun_out/r2z_smoke.log
timeout 1200 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "rc=$?"; tail -3 gpurun_out/r2z_bench.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2z_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')})
    print('roofline',d['roofline']['frac'], d['roofline']['achieved'], 'e2e',d['e2e']['value'])
    l=d['laplace3d']; print('laplace', l['value'], l['check']['equal'], {k:l['two_steps_per_pass'][k] for k in ('value','speedup_over_single_steps')}, l['two_steps_per_pass']['check']['equal_both_buffers'])
    for k,v in d['extra'].items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ('ms','GB/s','TFLOP/s','Gcells/s','frac_of_measured_hbm_peak','kernel','cublas_TFLOP/s')} if 'check' not in k else '')
except Exception as e: print('parse failed', e)
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err; echo "rc=$?"; cut -c1-300 gpurun_out/r2z_bench_ref.json
