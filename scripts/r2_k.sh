set -x
timeout 300 python profiles/check_fem_its.py > gpurun_out/r2k_fem_its.log 2>&1; tail -6 gpurun_out/r2k_fem_its.log
NG=8192 timeout 400 python profiles/time_dgemm.py b32x3 s32x3 s16x4 w32x3 w16x4 w16x3 > gpurun_out/r2k_dgemm.log 2>&1; grep -v "^{" gpurun_out/r2k_dgemm.log | tail -12
