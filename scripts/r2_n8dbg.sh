set -x
for dbg in 0 1 2 4 6 7; do
RACU_PUSH_DEBUG=$dbg timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2951$dbg bench.py --gpus 8 --workload stencil --steps 200 --warmup 10 > gpurun_out/r2_dbg$dbg.json 2> gpurun_out/r2_dbg$dbg.err
echo "dbg=$dbg $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_dbg$dbg.json | head -1)"
done
