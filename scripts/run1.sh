set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
RACU_NO_JIT=1 timeout 300 python profiles/time_generic.py > gpurun_out/generic_nojit.log 2>&1; cat gpurun_out/generic_nojit.log
timeout 300 python profiles/time_generic.py > gpurun_out/generic_jit.log 2>&1; cat gpurun_out/generic_jit.log
ls ra_ra_b200/jit_cache | wc -l
