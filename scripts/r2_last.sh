set -x
timeout 300 python -m pytest tests/test_gpu_stencil_steps.py -m gpu -x -q 2>&1 | tail -3
