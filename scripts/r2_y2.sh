set -x
timeout 300 python -m pytest tests/test_gpu_stencil_steps.py -m gpu -x -q 2>&1 | tail -2
CIS=61 timeout 300 python profiles/time_stencil2.py 2>&1 | tail -1 | tee gpurun_out/r2y2_stencil2.log | cut -c1-500
CIS=61 NS=768 F64=1 timeout 300 python profiles/time_stencil2.py 2>&1 | tail -1 | tee gpurun_out/r2y2_stencil2_f64.log | cut -c1-500
