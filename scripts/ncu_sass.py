"""Per-SASS-instruction view of an ncu report: executed count, samples and the dominant stall reason.
usage: python scripts/ncu_sass.py <report.ncu-rep> [min_exec_fraction_of_max=0.3]"""
import csv, subprocess, sys
rep = sys.argv[1]
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
col = {k: i for i, k in enumerate(hdr)}
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
data = rows[2:]
mx = max(int(r[col["Instructions Executed"]] or 0) for r in data)
tots = sum(int(r[col["# Samples"]] or 0) for r in data)
base = int(data[0][col["Address"]], 16)
for r in data:
    c = int(r[col["Instructions Executed"]] or 0)
    if c < frac * mx:
        continue
    s = int(r[col["# Samples"]] or 0)
    st = sorted(((int(r[col[k]] or 0), k[6:]) for k in stalls), reverse=True)[:2]
    print(f"{int(r[col['Address']],16)-base:6x} {c/1e6:8.1f}M {s/tots*100:5.2f}% {st[0][1]:>14}:{st[0][0]:<6} {st[1][1]:>12}:{st[1][0]:<6} {r[col['Source']][:70]}")
