# 8-GPU box: the default bench line (reduce weak scaling + laplace3d 1024^3 strong scaling) at N = 1, 2, 4, 8 and laplace3d 2048^3 at 8
set -x
run() { # N name args...
  N=$1; name=$2; shift 2
  if [ $N = 1 ]; then timeout 200 python bench.py --gpus 1 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  else timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; fi
  echo "rc=$?" >> gpurun_out/$name.err
}
run 8 r01d_n8 --extras 0 --e2e-steps 0
run 1 r01d_n1 --extras 0 --e2e-steps 0
run 4 r01d_n4 --extras 0 --e2e-steps 0
run 2 r01d_n2 --extras 0 --e2e-steps 0
run 8 r01d_stencil_8_2048 --workload stencil --steps 20 --warmup 4 --n3 2048
tail -2 gpurun_out/r01d_*.err
