REPS=5 python scripts/quick_perf.py C3 C3I | grep -E "rep[34]"
timeout 300 python -m pytest tests/test_gpu_grid.py tests/test_gpu_fullsize.py tests/test_gpu_edge.py -x -q -m gpu 2>&1 | tail -2
