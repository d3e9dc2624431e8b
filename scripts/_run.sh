timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/r1h_bench.json 2> gpurun_out/r1h_bench.err; cat gpurun_out/r1h_bench.json | python -c "import json,sys; d=json.load(sys.stdin); print({k:d[k] for k in ('value','ms_per_step','e2e','phases_ms')})"
