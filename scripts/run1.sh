set -x
timeout 300 python profiles/check_sgemm_tc5.py > gpurun_out/tc5_ts.log 2>&1; echo "rc=$?" >> gpurun_out/tc5_ts.log; tail -40 gpurun_out/tc5_ts.log
