set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
timeout 300 python profiles/time_generic.py > gpurun_out/generic_jit.log 2>&1; cat gpurun_out/generic_jit.log
