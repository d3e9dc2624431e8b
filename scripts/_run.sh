timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/r1j_bench.json 2> gpurun_out/r1j_bench.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1j_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1j_ncu_b.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_grid_nearest -s 1 -c 1 -o gpurun_out/r1j_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1j_ncu_full.log 2>&1
cat gpurun_out/r1j_bench.json | python -c "import json,sys; d=json.load(sys.stdin); print({k:d[k] for k in ('value','ms_per_step','phases_ms','gpu_launches')}, d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['cpu_baseline']['value'])"
