M2S_LIB=build/libm2s_stats.so python scripts/stats_pair.py 3
for p in 2 3; do
  M2S_PAIR=$p timeout 300 python -m pytest tests/test_gpu_grid.py tests/test_gpu_edge.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -2
done
M2S_PAIR=3 REPS=4 python scripts/quick_perf.py C3 C5 2>&1 | grep -E "rep[123]"
