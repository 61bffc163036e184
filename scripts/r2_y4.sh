set -x
RACU_ST2_PACKED=1 timeout 300 python -m pytest tests/test_gpu_stencil_steps.py -m gpu -x -q 2>&1 | tail -2
for pk in 0 1; do RACU_ST2_PACKED=$pk CIS=61 timeout 300 python profiles/time_stencil2.py 2>&1 | tail -1 | cut -c1-400 | sed "s/^/packed=$pk /"; done | tee gpurun_out/r2y4_packed.log
