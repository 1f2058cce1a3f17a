timeout 300 python tools/time_sample.py 2>&1 | tail -5
