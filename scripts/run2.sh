set -x

timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/multi2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/multi2.log
for halo in push nccl; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload stencil --steps 50 --warmup 5 --halo $halo > gpurun_out/st2_$halo.json 2> gpurun_out/st2_$halo.err; echo "rc=$?" >> gpurun_out/st2_$halo.err
done
tail -3 gpurun_out/multi2.log; cat gpurun_out/st2_push.json gpurun_out/st2_nccl.json; tail -5 gpurun_out/st2_push.err
