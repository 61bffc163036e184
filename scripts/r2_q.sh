set -x
timeout 600 python -m pytest tests/test_gpu_stencil_steps.py -m gpu -x -q > gpurun_out/r2q_pytest_steps.log 2>&1; echo "rc=$?" >> gpurun_out/r2q_pytest_steps.log; tail -3 gpurun_out/r2q_pytest_steps.log
for st in 4 5; do RACU_ST2_STAGES=$st CIS=31,61,91 timeout 300 python profiles/time_stencil2.py 2>&1 | grep "^[0-9]" | sed "s/^/stages=$st ci=/"; done | tee gpurun_out/r2q_stencil2_stages.log
NS=1024 TS=4 CIS=61 python profiles/time_stencil2.py > gpurun_out/r2q_plain.log 2>&1 &&
NS=1024 TS=4 CIS=61 ncu --set full --clock-control none --import-source on -k regex:stencil2 -c 1 -o gpurun_out/r2q_stencil2 python profiles/time_stencil2.py > gpurun_out/r2q_ncu.log 2>&1
tail -2 gpurun_out/r2q_ncu.log
