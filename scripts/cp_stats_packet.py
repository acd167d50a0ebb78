"""Debug: packet-search counters on the bench workload (FPOHM_CP_STATS=1): visits per warp, leaf steps, bail-outs."""
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
n = len(P) // 32 * 32
vis = (N[:n, 0] % 65536).reshape(-1, 32); lv = (N[:n, 0] // 65536).reshape(-1, 32)
code = N[:n, 2].reshape(-1, 32); cyc = N[:n, 1].reshape(-1, 32)[:, 0]
print("warps", len(vis), "| lanes settled by K1 %.4f, to K2 for tie-break %.4f, to K2 for search %.4f" % tuple((code == v).mean() for v in (0, 1, 2)))
print("warps that stopped sharing %.4f | warps with >=1 tie-break lane %.3f" % ((code == 2).any(1).mean(), (code == 1).any(1).mean()))
print("K1 node visits/warp mean %.1f p50 %.0f p99 %.0f max %.0f" % (vis[:, 0].mean(), np.median(vis[:, 0]), np.percentile(vis[:, 0], 99), vis[:, 0].max()))
print("K1 leaf steps/warp mean %.1f p99 %.0f" % (lv[:, 0].mean(), np.percentile(lv[:, 0], 99)))
print("K1 cycles/warp: mean %.0f p50 %.0f p99 %.0f max %.0f  sum %.3g | per step %.0f" % (cyc.mean(), np.median(cyc), np.percentile(cyc, 99), cyc.max(), cyc.sum(), (cyc / np.maximum(vis[:, 0] + lv[:, 0], 1)).mean()))
