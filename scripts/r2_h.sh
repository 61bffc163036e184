set -x
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2h_pytest_gpu.log; tail -8 gpurun_out/r2h_pytest_gpu.log
timeout 600 python profiles/time_shifted.py > gpurun_out/r2h_shifted.log 2>&1; grep -v "^{" gpurun_out/r2h_shifted.log | tail -20
