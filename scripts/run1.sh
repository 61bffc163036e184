set -x
timeout 600 python -m pytest tests/test_gpu_kat.py tests/test_cxx_header.py -m gpu -x -q > gpurun_out/pytest_kat.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_kat.log; tail -25 gpurun_out/pytest_kat.log
