set -x
timeout 900 python -m pytest tests/test_gpu_stencil_steps.py tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_cxx_header.py -m gpu -x -q > gpurun_out/r2z2_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2z2_pytest.log; tail -3 gpurun_out/r2z2_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python bench.py > gpurun_out/r2z2_bench.json 2> gpurun_out/r2z2_bench.err; echo "rc=$?"; tail -2 gpurun_out/r2z2_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z2_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'roofline',d['roofline']['frac'], 'e2e',d['e2e']['value'])
l=d['laplace3d']; print('laplace', l['value'], l['check']['equal'], {k:l['two_steps_per_pass'][k] for k in ('value','speedup_over_single_steps')}, l['two_steps_per_pass']['check']['equal_both_buffers'])
PY
