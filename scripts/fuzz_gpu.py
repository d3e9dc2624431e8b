"""GPU fuzz: random meshes x random grids / queries, every result compared with the exact oracle (the cases of
tests/fuzz_cases.py; `pytest -m gpu` runs 200 of them). usage: python scripts/fuzz_gpu.py [n_cases=150] [seed=0]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mesh_to_sdf_b200 as m2s
from fuzz_cases import run_fuzz

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 150
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
t0 = time.time()
bad = run_fuzz(m2s.Context(), n_cases, seed, log=lambda *a: print(*a, flush=True))
print(f"fuzz: {n_cases} cases, {bad} mismatches, {time.time() - t0:.1f} s")
sys.exit(1 if bad else 0)
