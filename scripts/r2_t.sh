set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kat.py -m gpu -x -q -k "mass or 1d3 or shifted or stencil or kat" > gpurun_out/r2t_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2t_pytest.log; tail -5 gpurun_out/r2t_pytest.log
NSH=4100,8196 timeout 600 python profiles/time_shifted.py > gpurun_out/r2t_shifted.log 2>&1; grep -v "^{" gpurun_out/r2t_shifted.log | tail -20
for ci in 2 4 16; do RACU_ST_CI2=$ci NSH=8196 timeout 600 python profiles/time_shifted.py 2>&1 | grep "stiff\|mass" | sed "s/^/ci2=$ci /"; done
