set -x
timeout 600 python -m pytest tests/test_gpu_stencil_steps.py -m gpu -x -q > gpurun_out/r2o_pytest_steps.log 2>&1; echo "rc=$?" >> gpurun_out/r2o_pytest_steps.log; tail -5 gpurun_out/r2o_pytest_steps.log
CIS=16,31,46,61,91 timeout 300 python profiles/time_stencil2.py > gpurun_out/r2o_stencil2_ci.log 2>&1; tail -8 gpurun_out/r2o_stencil2_ci.log | cut -c1-300
NS=512 TS=4 CIS=61 python profiles/time_stencil2.py > gpurun_out/r2o_plain.log 2>&1 &&
NS=512 TS=4 CIS=61 ncu --set full --clock-control none --import-source on -k regex:stencil2 -c 1 -o gpurun_out/r2o_stencil2 python profiles/time_stencil2.py > gpurun_out/r2o_ncu.log 2>&1
tail -3 gpurun_out/r2o_ncu.log
