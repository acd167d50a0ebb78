"""Debug: traversal counters of the closest-point kernel on the bench workload (FPOHM_CP_STATS=1)."""
import os, sys
from pathlib import Path
os.environ["FPOHM_CP_STATS"] = "1"
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import fpohm_b200 as fp
import bench
ctx = fp.Context(0)
V, F = bench.workload(fp)
mesh = fp.TriMesh(ctx, V, F)
prm = fp.octree_grid_setup(V, 1 << 20); prm.c.stop_extent = 1 << bench.STOP_E
o = fp.Octree.build(ctx, mesh, prm)
Vh, H, _ = o.hexes()
ext = Vh[H[:, 1].astype(np.int64), 0] - Vh[H[:, 0].astype(np.int64), 0]
P = bench.make_queries(Vh, H, ext)
S, I, C, N = mesh.signed_distance_pseudonormal(P)
nodes, leaves = N[:, 0], N[:, 1]
print("queries", len(P), "nodes/query mean %.1f median %.0f p99 %.0f max %.0f" % (nodes.mean(), np.median(nodes), np.percentile(nodes, 99), nodes.max()))
print("leaves/query mean %.1f median %.0f p99 %.0f max %.0f" % (leaves.mean(), np.median(leaves), np.percentile(leaves, 99), leaves.max()))
w = nodes.reshape(-1, 32) if len(nodes) % 32 == 0 else nodes[: len(nodes) // 32 * 32].reshape(-1, 32)
print("per-warp max/mean node visits: %.2f" % (w.max(1).mean() / w.mean()))
d = np.abs(S); print("distance/leaf-extent median %.2f" % np.median(d / np.repeat(ext, 4)))
srt = np.sort(nodes)[::-1]
for frac in (1e-5, 1e-4, 1e-3, 1e-2, 1e-1):
    k = max(1, int(len(srt) * frac))
    print("top %.3f%% of queries hold %.1f%% of the node visits (min visits in that set %.0f)" % (100 * frac, 100 * srt[:k].sum() / srt.sum(), srt[k - 1]))
wm = w.max(1)
print("per-warp max visits: mean %.0f p99 %.0f p99.9 %.0f max %.0f" % (wm.mean(), np.percentile(wm, 99), np.percentile(wm, 99.9), wm.max()))
