timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "ln_modulate" 2>&1 | tail -2
timeout 300 python tools/time_ln.py 2>&1 | tail -4
