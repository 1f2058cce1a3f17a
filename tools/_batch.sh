CVAR_CONV3_ROWS4=0 timeout 200 python tools/time_conv_out.py 2>&1 | tail -1
timeout 200 python tools/time_conv_out.py 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "conv_out or fhat" 2>&1 | tail -3
