timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline | python -c "import json,sys; d=json.load(sys.stdin); print({k:d[k] for k in ('value','ms_per_step','phases_ms')}, d['e2e']['value'], d['e2e']['ms_per_step'])"
python scripts/quick_perf.py C4 C2 | grep rep2
