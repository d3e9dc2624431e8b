"""Turns gpurun_out/ ncu artefacts into the tracked summaries under profiles/.
usage: python scripts/profile_summary.py <tag> <launches.csv> <full.ncu-rep> <kernel-substring>"""
import collections, csv, json, os, re, subprocess, sys
tag, launches, rep, kern = sys.argv[1:5]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prof = os.path.join(root, "profiles")
os.makedirs(prof, exist_ok=True)

# ---- launch list -> per-kernel share ----
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr = rows[0]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
per = collections.OrderedDict()
seq = []
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").strip()
    name = re.sub(r"m2s::<unnamed>::|m2s::\(anonymous namespace\)::", "", name)
    per.setdefault(name, []).append(v * scale)
    seq.append((name, v * scale))
total = sum(sum(v) for v in per.values())
with open(os.path.join(prof, f"{tag}_launches_summary.md"), "w") as f:
    f.write(f"# ncu launch list summary ({tag})\n\nSource: `ncu --metrics gpu__time_duration.sum --clock-control none` "
            f"around `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (cold-cache, serialised: compare SHARES).\n"
            f"All launches of the process are listed (warm-up, timed and e2e steps alike).\n\n"
            f"| kernel | launches | total us | mean us | share |\n|---|---:|---:|---:|---:|\n")
    for name, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"| `{name[:90]}` | {len(v)} | {sum(v):.1f} | {sum(v)/len(v):.1f} | {sum(v)/total*100:.2f}% |\n")
    f.write(f"\nTotal device time in the list: {total/1e3:.2f} ms over {len(seq)} launches.\n")
import shutil
shutil.copy(launches, os.path.join(prof, f"{tag}_launches.csv"))

# ---- full capture -> key metrics ----
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
h, u, v = r[0], r[1], r[2]
m = {a: (c, b) for a, b, c in zip(h, u, v)}
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.avg.per_second"]
def num(k):
    try: return float(m[k][0].replace(",", ""))
    except Exception: return None
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rd = num("dram__bytes_read.sum") * scale.get(m["dram__bytes_read.sum"][1], 1)
wr = num("dram__bytes_write.sum") * scale.get(m["dram__bytes_write.sum"][1], 1)
lines = subprocess.run([sys.executable, os.path.join(root, "scripts", "ncu_lines.py"), rep, kern, "30"], capture_output=True, text=True).stdout
with open(os.path.join(prof, f"{tag}_{kern.split('IL')[0]}_ncu_full.md"), "w") as f:
    f.write(f"# ncu --set full: {r[2][h.index('Kernel Name')][:160] if 'Kernel Name' in h else kern} ({tag})\n\n"
            f"Captured with `ncu --set full --clock-control none --import-source on -k regex:{kern} -c 1` around "
            f"`{os.environ.get('PROFILE_CMD', 'python scripts/ncu_case.py')}` ({os.environ.get('PROFILE_DESC', 'workload C3, the whole 256^3 grid in one launch')}). "
            f"Numbers under a profiler are not bench values.\n\n| metric | value | unit |\n|---|---:|---|\n")
    for k in keys:
        if k in m:
            f.write(f"| `{k}` | {m[k][0]} | {m[k][1]} |\n")
    f.write(f"\nDRAM traffic per launch: read {rd/1e6:.1f} MB + write {wr/1e6:.1f} MB = {(rd+wr)/1e6:.1f} MB "
            f"({os.environ.get('PROFILE_ALG', 'algorithmic bytes: 68.9 MB = 64 MiB output + 1.8 MB mesh')}).\n\n## Hot source lines (SASS joined with -lineinfo)\n\n```\n{lines}```\n")
# bench.py reads roofline.traffic from here: one key per (kernel, workload)
tj = os.path.join(prof, "roofline_traffic.json")
try:
    traffic = json.load(open(tj))
except Exception:
    traffic = {}
traffic[os.environ.get("PROFILE_KEY", "k_grid_nearest_dram_bytes_per_launch_C3")] = rd + wr
traffic.setdefault("sources", {})[os.environ.get("PROFILE_KEY", "k_grid_nearest_dram_bytes_per_launch_C3")] = f"profiles/{tag}_{kern.split('IL')[0]}_ncu_full.md"
json.dump(traffic, open(tj, "w"), indent=1)
print("ok", rd, wr)
