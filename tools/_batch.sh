timeout 300 python tools/trace_conv.py 2>&1 | tail -8
