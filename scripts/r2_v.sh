set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_jit.py -m gpu -x -q -k "staged or shifted or mass or 1d3 or jit" > gpurun_out/r2v_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2v_pytest.log; tail -12 gpurun_out/r2v_pytest.log
NSH=4097,8193 timeout 600 python profiles/time_shifted.py > gpurun_out/r2v_shifted.log 2>&1; grep -v "^{" gpurun_out/r2v_shifted.log | tail -20
RACU_NO_STAGED=1 NSH=8193 timeout 600 python profiles/time_shifted.py 2>&1 | grep -v "^{" | grep "8193" | sed "s/^/nostaged /"
