set -x
timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/multi4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/multi4.log; tail -3 gpurun_out/multi4.log
run() { name=$1; shift
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --workload stencil --n0 512 --steps 200 --warmup 10 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "rc=$?" >> gpurun_out/$name.err; }
run r01f_push
RACU_PUSH_DEBUG=7 run r01f_none
for f in gpurun_out/r01f_*.json; do echo $f; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' $f | head -2; done
