set -x
timeout 300 python profiles/check_sgemm_tc5.py > gpurun_out/tc5.log 2>&1; echo "rc=$?" >> gpurun_out/tc5.log; tail -60 gpurun_out/tc5.log
RACU_JIT_VERBOSE=1 timeout 200 python profiles/time_folds_percall.py > gpurun_out/folds_percall.log 2>&1; cat gpurun_out/folds_percall.log
