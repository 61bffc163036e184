set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scatter" > gpurun_out/r2u_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2u_pytest.log; tail -3 gpurun_out/r2u_pytest.log
timeout 300 python profiles/time_fem.py > gpurun_out/r2u_fem.log 2>&1; tail -1 gpurun_out/r2u_fem.log
for band in 256 1024 2048; do RACU_ST_BAND=$band CIS=61 timeout 300 python profiles/time_stencil2.py 2>&1 | grep "^61" | sed "s/^/band=$band ci=/"; done | tee gpurun_out/r2u_stencil2_band.log
F64=1 NS=768 CIS=61 timeout 300 python profiles/time_stencil2.py 2>&1 | tail -1 | tee gpurun_out/r2u_stencil2_f64.log | cut -c1-500
