timeout 300 python -m pytest tests/test_gpu_grid.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -2
for c in -1 0 12 25 40; do echo "== carveout $c"; M2S_CARVEOUT=$c REPS=4 python scripts/quick_perf.py C3 2>&1 | grep -E "rep[3]"; done
