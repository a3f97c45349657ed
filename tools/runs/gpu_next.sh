mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_next.py -m gpu -x -q > gpurun_out/pytest_next.log 2>&1; tail -15 gpurun_out/pytest_next.log
