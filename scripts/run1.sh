set -x
for b in 100000000 512 256 1024; do RACU_ST_BAND=$b timeout 120 python profiles/time_stencil_shapes.py >> gpurun_out/bands.log 2>gpurun_out/bands.err; done
RACU_ST_CI=8 RACU_ST_BAND=100000000 timeout 120 python profiles/time_stencil_shapes.py >> gpurun_out/bands.log 2>>gpurun_out/bands.err
RACU_ST_CI=8 RACU_ST_BAND=512 timeout 120 python profiles/time_stencil_shapes.py >> gpurun_out/bands.log 2>>gpurun_out/bands.err
cat gpurun_out/bands.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"; cut -c1-600 gpurun_out/bench_default.json
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2>&1; cat gpurun_out/bench_reference.json | cut -c1-400
