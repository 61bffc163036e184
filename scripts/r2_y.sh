set -x
timeout 300 python -m pytest tests/test_gpu_stencil_steps.py -m gpu -x -q 2>&1 | tail -2
CIS=61 timeout 300 python profiles/time_stencil2.py 2>&1 | tail -1 | tee gpurun_out/r2y_stencil2.log | cut -c1-500
NS=1024 TS=4 CIS=61 python profiles/time_stencil2.py > gpurun_out/r2y_plain.log 2>&1 &&
NS=1024 TS=4 CIS=61 ncu --set full --clock-control none --import-source on -k regex:stencil2 -c 1 -o gpurun_out/r2y_stencil2 python profiles/time_stencil2.py > gpurun_out/r2y_ncu.log 2>&1
tail -1 gpurun_out/r2y_ncu.log
