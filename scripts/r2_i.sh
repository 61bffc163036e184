set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cxx_header.py tests/test_gpu_kat.py -m gpu -x -q > gpurun_out/r2i_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_pytest_gpu.log; tail -5 gpurun_out/r2i_pytest_gpu.log
timeout 600 python profiles/time_shifted.py > gpurun_out/r2i_shifted.log 2>&1; grep -v "^{" gpurun_out/r2i_shifted.log | tail -20
for v in 16x3 16x4 32x3; do RACU_DGEMM=$v NG=8192 timeout 300 python profiles/time_gemm.py 2>&1 | grep "gemm f64" | sed "s/^/$v /"; done | tee gpurun_out/r2i_dgemm.log
