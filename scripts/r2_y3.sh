set -x
NS=1024 TS=4 CIS=61 python profiles/time_stencil2.py > gpurun_out/r2y3_plain.log 2>&1 &&
NS=1024 TS=4 CIS=61 ncu --set full --clock-control none --import-source on -k regex:stencil2 -c 1 -o gpurun_out/r2y3_stencil2 python profiles/time_stencil2.py > gpurun_out/r2y3_ncu.log 2>&1
tail -1 gpurun_out/r2y3_ncu.log
