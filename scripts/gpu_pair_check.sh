#!/bin/bash
# development aid: parity + timing of the pair kernel variants (M2S_PAIR = 0 off, 2 x-extended, 3 z-extended)
mkdir -p gpurun_out
for p in 2 3; do
  M2S_PAIR=$p timeout 300 python -m pytest tests/test_gpu_grid.py tests/test_gpu_edge.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -4
done
for p in 0 2 3; do
  echo "== M2S_PAIR=$p"; M2S_PAIR=$p REPS=4 timeout 120 python scripts/quick_perf.py C3 C5 2>&1 | grep -E "rep[23]|neg"
done
