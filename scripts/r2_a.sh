set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_pytest_gpu.log; tail -15 gpurun_out/r2a_pytest_gpu.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; tail -1 gpurun_out/r2a_smoke.log
timeout 600 python profiles/time_folds.py > gpurun_out/r2a_folds.log 2>&1; tail -60 gpurun_out/r2a_folds.log
