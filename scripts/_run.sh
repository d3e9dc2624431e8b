timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu --durations=5 2>&1 | tail -12
