set -x
P="python profiles/prof_ply.py gemv32 gevm32 gevm64 sgemm"
$P > gpurun_out/plain_prof2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'fold|tc5' -c 12 -o gpurun_out/r01c_gemv_full $P > gpurun_out/ncu_f2.log 2>&1
tail -3 gpurun_out/ncu_f2.log; cat gpurun_out/plain_prof2.log | tail -5
