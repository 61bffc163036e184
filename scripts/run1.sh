set -x
timeout 300 python profiles/check_sgemm_tc5.py > gpurun_out/tc5.log 2>&1; echo "rc=$?" >> gpurun_out/tc5.log; tail -12 gpurun_out/tc5.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gemm" 2>&1 | tail -3
