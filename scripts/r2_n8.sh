set -x
timeout 200 python bench.py --gpus 1 --workload stencil --steps 100 --warmup 10 > gpurun_out/r2_stencil_n1.json 2> gpurun_out/r2_stencil_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --workload stencil --steps 200 --warmup 10 > gpurun_out/r2_stencil_n8.json 2> gpurun_out/r2_stencil_n8.err
grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"halo_timeouts": [0-9a-z]*' gpurun_out/r2_stencil_n1.json gpurun_out/r2_stencil_n8.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "rc=$?"
tail -3 gpurun_out/r2_bench_n8.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_n8.json').read().strip().splitlines() if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, d['e2e']['value'])
l=d['laplace3d']; print({k:l[k] for k in ('value','ms_per_step','halo_timeouts','check')})
l=d.get('laplace3d_weak'); print(l and {k:l[k] for k in ('value','ms_per_step','halo_timeouts')})
print(d.get('map_sharded'))
PY
