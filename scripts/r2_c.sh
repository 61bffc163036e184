set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_pytest_gpu.log; tail -5 gpurun_out/r2c_pytest_gpu.log
timeout 600 python profiles/time_folds.py > gpurun_out/r2c_folds.log 2>&1; grep -v "^{" gpurun_out/r2c_folds.log | tail -40
timeout 600 tests/cxx/build/bench_header_cuda 28 5 > gpurun_out/r2c_bench_header.json 2> gpurun_out/r2c_bench_header.err; cat gpurun_out/r2c_bench_header.json; tail -3 gpurun_out/r2c_bench_header.err
