set -x
timeout 100 python bench.py --gpus 1 --workload stencil --steps 100 --warmup 10 > gpurun_out/r01g_stencil_1.json 2> gpurun_out/r01g_stencil_1.err
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --workload stencil --steps 200 --warmup 10 > gpurun_out/r01g_stencil_8.json 2> gpurun_out/r01g_stencil_8.err
grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"halo_timeouts": [0-9a-z]*' gpurun_out/r01g_stencil_1.json gpurun_out/r01g_stencil_8.json
