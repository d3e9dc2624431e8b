import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["M2S_STATS"] = "2"; os.environ["M2S_DYNAMIC"] = "0"
import numpy as np
import mesh_to_sdf_b200 as m2s
from mesh_to_sdf_b200 import synth
verts, tris = synth.bumpy_torus(256, 196)
mn, mx = synth.padded_grid_box(verts)
n = 256
grid = m2s.Grid.from_bounding_box(mn, mx, [n, n, n])
for leaf in (4,):
    os.environ["M2S_LEAF_SIZE"] = str(leaf)
    with m2s.Context() as ctx:
        w = ctx.grid_sdf(verts, tris, grid, 0).view(np.uint32)
    nodes = (w & 0xffff).astype(np.int64); leaves = (w >> 16).astype(np.int64)
    print("leaf", leaf, "mean nodes", nodes.mean(), "leaves", leaves.mean())
    qs = [0.1, 0.25, 0.5, 0.75, 0.9, 0.95, 0.99, 0.999, 1.0]
    print(" node quantiles", dict(zip(qs, np.quantile(nodes, qs))))
    print(" leaf quantiles", dict(zip(qs, np.quantile(leaves, qs))))
    # share of total work in the top x% voxels
    srt = np.sort(nodes.ravel())[::-1]; cs = np.cumsum(srt) / srt.sum()
    for f in (0.01, 0.05, 0.1, 0.25, 0.5):
        print(f"  top {f*100:.0f}% voxels hold {cs[int(f*len(srt))-1]*100:.1f}% of node visits")
    # per-warp max vs mean: warps are 2x4x4 tiles
    t = nodes.reshape(n//2, 2, n//4, 4, n//4, 4).transpose(0,2,4,1,3,5).reshape(-1, 32)
    print("  warp efficiency (sum / (32*max)) for total nodes:", t.sum() / (32 * t.max(axis=1).sum()))
    np.save("gpurun_out/workmap_nodes.npy", nodes.astype(np.uint16).reshape(n, n, n)[::4, ::4, ::4])
