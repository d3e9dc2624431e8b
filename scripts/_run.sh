M2S_LIB=build/libm2s_stats.so python scripts/stats_pair.py 0 2 3
for v in mb4 mb6; do echo "== $v"; M2S_LIB=build/libm2s_$v.so M2S_PAIR=3 REPS=4 python scripts/quick_perf.py C3 2>&1 | grep -E "rep[23]"; done
M2S_PAIR=3 timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_grid_nearest -s 1 -c 1 -o gpurun_out/r1f_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1f_ncu_full.log 2>&1
tail -2 gpurun_out/r1f_ncu_full.log
