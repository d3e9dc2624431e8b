timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
REPS=4 python scripts/quick_perf.py C3 C2 C4 | grep -E "rep[23]"
python scripts/_c5.py | tail -1
timeout 300 python scripts/fuzz_gpu.py 300 11 | tail -3
