timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/r1g_bench.json 2> gpurun_out/r1g_bench.err; cat gpurun_out/r1g_bench.json
python scripts/quick_perf.py C2 C4 C5 2>&1 | grep -E "rep[12]"
