"""Joins an ncu report's SASS-level instruction statistics with source lines (nvdisasm -g of the in-tree cubin).
usage: python scripts/ncu_lines.py gpurun_out/prof.ncu-rep <kernel-name-substring> [top_n]"""
import collections, csv, os, re, subprocess, sys, tempfile
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
lib = os.environ.get("M2S_LIB", os.path.join(root, "mesh_to_sdf_b200", "libm2s.so"))
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
sass = "".join(subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
               for f in sorted(os.listdir(tmp)) if f.endswith(".cubin"))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
kname = rows[0][1]
print("kernel:", kname[:120])
# mangled-name match: pick the function whose instruction count equals the report's
hdr = rows[1]
ai, ci, ti, si = (hdr.index(k) for k in ("Address", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
data = [(int(r[ai], 16), int(r[ci] or 0), int(r[ti] or 0), int(r[si] or 0)) for r in rows[2:] if len(r) > ci]
base = data[0][0]
fns = collections.OrderedDict()
cur_fn, cur_line = None, None
for l in sass.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        cur_fn = m.group(1); fns[cur_fn] = {}; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and cur_fn:
        fns[cur_fn][int(m.group(1), 16)] = cur_line
cands = [f for f in fns if kern in f and len(fns[f]) == len(data)]
if not cands:
    cands = [f for f in fns if kern in f]
    print("warning: no exact size match; candidates", [(f[-60:], len(fns[f])) for f in cands], "report has", len(data))
amap = fns[cands[0]]
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = tott = tots = 0
for a, c, t, s in data:
    ln = amap.get(a - base)
    agg[ln][0] += c; agg[ln][1] += t; agg[ln][2] += s
    tot += c; tott += t; tots += s
print(f"warp instr {tot:.4g}  thread instr {tott:.4g}  avg active threads {tott / tot:.2f}")
src = {}
for name in os.listdir(os.path.join(root, "mesh_to_sdf_b200", "csrc")):
    src[name] = open(os.path.join(root, "mesh_to_sdf_b200", "csrc", name)).read().splitlines()
for ln, (c, t, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    if ln and ln[0] in src and ln[1] - 1 < len(src[ln[0]]):
        text = src[ln[0]][ln[1] - 1].strip()[:100]
    print(f"{c / tot * 100:5.1f}% inst {s / max(tots, 1) * 100:5.1f}% smp  thr/inst {t / max(c, 1):5.1f}  {ln}  {text}")
