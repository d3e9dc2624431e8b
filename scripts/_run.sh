for p in 1 8; do
  M2S_PAIR=$p timeout 400 python -m pytest tests/test_gpu_points.py tests/test_gpu_edge.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -3
done
for p in 0 1 8; do echo "== M2S_PAIR=$p"; M2S_PAIR=$p python scripts/quick_perf.py C4 2>&1 | grep -E "rep[2]"; done
