set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_jit.py tests/test_gpu_kat.py -m gpu -x -q > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_pytest_gpu.log; tail -5 gpurun_out/r2b_pytest_gpu.log
timeout 600 python profiles/time_folds.py > gpurun_out/r2b_folds.log 2>&1; grep -v "^{" gpurun_out/r2b_folds.log | tail -60
