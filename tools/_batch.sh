mkdir -p gpurun_out
timeout 300 python tools/time_sample.py 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_sampler.py tests/test_gpu_conditional.py -m gpu -x -q 2>&1 | tail -3
