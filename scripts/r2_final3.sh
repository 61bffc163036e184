set -x
timeout 1800 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2z3_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2z3_pytest_gpu.log; tail -14 gpurun_out/r2z3_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
