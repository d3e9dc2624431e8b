timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python scripts/quick_perf.py C2 C3N 2>&1 | grep -E "rep[12]"
M2S_PAIR=0 python scripts/quick_perf.py C2 2>&1 | grep -E "rep[12]"
