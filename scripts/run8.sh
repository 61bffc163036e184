# 8-GPU box: multi-GPU parity (4 ranks: middle ranks have two neighbours) + strong scaling of the stencil, weak of the folds
set -x
timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/multi8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/multi8.log
run() { # N name args...
  N=$1; name=$2; shift 2
  if [ $N = 1 ]; then timeout 150 python bench.py --gpus 1 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  else timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; fi
  echo "rc=$?" >> gpurun_out/$name.err
}
run 1 r01b_stencil_1 --workload stencil --steps 100 --warmup 10
run 8 r01b_stencil_8 --workload stencil --steps 100 --warmup 10
run 4 r01b_stencil_4 --workload stencil --steps 100 --warmup 10
run 2 r01b_stencil_2 --workload stencil --steps 100 --warmup 10
run 8 r01b_stencil_8_2048 --workload stencil --steps 20 --warmup 4 --n3 2048
run 8 r01b_reduce_8 --steps 20 --warmup 3 --e2e-steps 0
tail -3 gpurun_out/multi8.log; cat gpurun_out/r01b_*.json | cut -c1-330
