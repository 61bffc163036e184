set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > gpurun_out/r01h_default_n2.json 2> gpurun_out/r01h_default_n2.err; echo "rc=$?"
tail -3 gpurun_out/r01h_default_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r01h_default_n2.json').read().strip().splitlines() if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, d['e2e'], d['laplace3d'] and {k:d['laplace3d'][k] for k in ('value','ms_per_step','halo_timeouts')})
PY
