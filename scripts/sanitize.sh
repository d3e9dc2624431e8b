#!/bin/bash
# compute-sanitizer over the shipped kernels: memcheck, racecheck, synccheck, initcheck on a 64^3
# Raycast + Normal grid, a pipelined pageable copy, a 50k-query C4-shaped call and the render order of the grid (radix sort). Summaries -> gpurun_out/sanitize_*.log
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/m2s_sanitize_case.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
import mesh_to_sdf_b200 as m2s
from mesh_to_sdf_b200 import synth
verts, tris = synth.bumpy_torus(64, 40)
mn, mx = synth.padded_grid_box(verts)
grid = m2s.Grid.from_bounding_box(mn, mx, [64, 64, 64])
q = synth.splitmix64_points(50000, mn, mx)
with m2s.Context() as c:
    a = c.grid_sdf(verts, tris, grid, 0)
    b = c.grid_sdf(verts, tris, grid, 1)
    c.set_option(m2s.OPT_HOST_PATH, m2s.HOST_PIPELINED)
    a2 = c.grid_sdf(verts, tris, grid, 0)
    assert np.array_equal(a.view(np.uint32), a2.view(np.uint32))
    for accel, sign in [(3, 0), (0, 0), (1, 1), (2, 0)]:
        c.sdf(verts, tris, q, accel, sign)
    with c.mesh(verts, tris) as mesh:
        mesh.grid_sdf(grid, 0)
        mesh.sdf(q[:5000], 3)
    order, (lo, hi) = c.grid_order(a)  # the radix sort over 64 tiles (look-back), iso limits on its histogram read
    assert np.all(np.diff(a[order]) >= 0) and lo == a.min() and hi == a.max()
print("case ok", float(np.abs(a).sum()), float(np.abs(b).sum()))
PY
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/m2s_sanitize_case.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|case ok' gpurun_out/sanitize_$tool.log | tr '\n' ' ')"
done
