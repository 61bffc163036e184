set -x
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2b_n2_pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_n2_pytest_multi.log; tail -4 gpurun_out/r2b_n2_pytest_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --workload stencil --steps 100 --warmup 6 > gpurun_out/r2b_stencil_n2.json 2> gpurun_out/r2b_stencil_n2.err; echo "rc=$?"
tail -2 gpurun_out/r2b_stencil_n2.err; cut -c1-400 gpurun_out/r2b_stencil_n2.json
timeout 300 python profiles/time_fem.py > gpurun_out/r2b_fem.log 2>&1; tail -1 gpurun_out/r2b_fem.log
