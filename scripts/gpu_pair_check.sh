#!/bin/bash
# development aid: parity + timing of the multi-voxel-per-lane grid kernels (M2S_PAIR, see launch_grid_final)
VARIANTS=${VARIANTS:-"4 5 6 7"}
for p in $VARIANTS; do
  M2S_PAIR=$p timeout 300 python -m pytest tests/test_gpu_grid.py tests/test_gpu_edge.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -2
done
for p in ${TIMED:-"3 $VARIANTS"}; do
  echo "== M2S_PAIR=$p"; M2S_PAIR=$p REPS=4 timeout 120 python scripts/quick_perf.py C3 2>&1 | grep -E "rep[23]"
done
