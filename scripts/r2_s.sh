set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mass or 1d3 or shifted or stencil" > gpurun_out/r2s_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2s_pytest.log; tail -12 gpurun_out/r2s_pytest.log
NSH=4100,8196 timeout 600 python profiles/time_shifted.py > gpurun_out/r2s_shifted.log 2>&1; grep -v "^{" gpurun_out/r2s_shifted.log | tail -20
