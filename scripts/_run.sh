python scripts/_c5.py | tail -1
REPS=3 python scripts/quick_perf.py C3 | grep rep2
timeout 300 python -m pytest tests/test_gpu_grid.py tests/test_gpu_points.py -x -q -m gpu 2>&1 | tail -2
