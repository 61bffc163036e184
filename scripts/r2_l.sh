set -x
timeout 600 python -m pytest tests/test_gpu_stencil_steps.py -m gpu -x -q > gpurun_out/r2l_pytest_steps.log 2>&1; echo "rc=$?" >> gpurun_out/r2l_pytest_steps.log; tail -15 gpurun_out/r2l_pytest_steps.log
CIS=4,8,12,16 timeout 300 python profiles/time_stencil2.py > gpurun_out/r2l_stencil2.log 2>&1; tail -8 gpurun_out/r2l_stencil2.log
NG=16384 timeout 300 python profiles/time_dgemm.py b32x3 w32x3 w16x4 > gpurun_out/r2l_dgemm16k.log 2>&1; grep -v "^{" gpurun_out/r2l_dgemm16k.log | tail
