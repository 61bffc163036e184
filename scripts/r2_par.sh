set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_jit.py -m gpu -x -q --durations=5 > gpurun_out/r2p2_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2p2_pytest.log; tail -10 gpurun_out/r2p2_pytest.log
ls ra_ra_b200/jit_cache | wc -l
