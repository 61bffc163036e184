set -x
timeout 1800 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_pytest_gpu.log; tail -22 gpurun_out/r2f_pytest_gpu.log
timeout 600 python profiles/time_shifted.py > gpurun_out/r2f_shifted.log 2>&1; grep -v "^{" gpurun_out/r2f_shifted.log | tail -20
timeout 600 python bench.py --steps 5 --warmup 3 --laplace 0 --e2e-steps 0 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'])
for k,v in d['extra'].items():
    if k.startswith('map'): print(k, v)
PY
