set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_jit.py -m gpu -x -q > gpurun_out/r2w_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2w_pytest.log; tail -5 gpurun_out/r2w_pytest.log
NSH=4097,8193 timeout 600 python profiles/time_shifted.py > gpurun_out/r2w_shifted.log 2>&1; grep -v "^{" gpurun_out/r2w_shifted.log | tail -20
