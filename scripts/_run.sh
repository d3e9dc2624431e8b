for v in "" build/libm2s_w8.so; do echo "== lib=$v"; M2S_LIB=$v REPS=4 python scripts/quick_perf.py C3 C3I | grep -E "rep3"; done
M2S_LIB=build/libm2s_w8.so timeout 300 python -m pytest tests/test_gpu_grid.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -2
