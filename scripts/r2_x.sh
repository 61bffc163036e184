set -x
for st in 3 5; do RACU_ST2_STAGES=$st CIS=61 timeout 300 python profiles/time_stencil2.py 2>&1 | grep "^61" | sed "s/^/stages=$st ci=/"; done | tee gpurun_out/r2x_stencil2_occ.log
timeout 300 python -m pytest tests/test_gpu_stencil_steps.py -m gpu -x -q 2>&1 | tail -2
RACU_ST2_STAGES=3 timeout 300 python -m pytest tests/test_gpu_stencil_steps.py -m gpu -x -q 2>&1 | tail -2
