set -x
CIS=16,24,32,64 timeout 300 python profiles/time_stencil2.py > gpurun_out/r2m_stencil2_ci.log 2>&1; tail -5 gpurun_out/r2m_stencil2_ci.log | cut -c1-400
for band in 128 256 1024; do RACU_ST_BAND=$band CIS=16 timeout 300 python profiles/time_stencil2.py 2>&1 | tail -1 | cut -c1-600 | sed "s/^/band=$band /"; done | tee gpurun_out/r2m_stencil2_band.log
NS=512 TS=4 CIS=16 python profiles/time_stencil2.py > gpurun_out/r2m_plain.log 2>&1 &&
NS=512 TS=4 CIS=16 ncu --set full --clock-control none --import-source on -k regex:stencil2 -c 1 -o gpurun_out/r2m_stencil2 python profiles/time_stencil2.py > gpurun_out/r2m_ncu.log 2>&1
tail -3 gpurun_out/r2m_ncu.log
