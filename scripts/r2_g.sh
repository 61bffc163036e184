set -x
timeout 1800 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r2g_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_pytest_gpu.log; tail -14 gpurun_out/r2g_pytest_gpu.log
timeout 600 python profiles/time_shifted.py > gpurun_out/r2g_shifted.log 2>&1; grep -v "^{" gpurun_out/r2g_shifted.log | tail -20
