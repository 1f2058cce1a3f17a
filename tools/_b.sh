timeout 900 python -m pytest tests/test_gpu_conditional.py -m gpu -x -q -k "forward" 2>&1 | tail -4
