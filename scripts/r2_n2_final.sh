set -x
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2e_n2_pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_n2_pytest_multi.log; tail -6 gpurun_out/r2e_n2_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err; echo "rc=$?"
tail -3 gpurun_out/r2e_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2e_bench_n2.json').read().strip().splitlines() if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, d['e2e']['value'])
l=d['laplace3d']; print({k:l[k] for k in ('value','ms_per_step','halo_timeouts','check')})
print(d.get('map_sharded'))
PY
