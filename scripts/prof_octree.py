"""Profiling driver: mesh upload + N octree builds (+ hexes) on the bench workload; run under ncu time-only metrics."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import fpohm_b200 as fp

which = sys.argv[1] if len(sys.argv) > 1 else "gear"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 12
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
pm = fp.procedural
if which == "gear":
    V, F, _ = pm.gear()
elif which == "torus":
    V, F = pm.torus()
else:
    V, F = pm.linked_tori(4, 64, 32); V, F = pm.midpoint_subdivide(V, F, int(which[-1]) if which[-1].isdigit() else 1)
ctx = fp.Context(0)
m = fp.TriMesh(ctx, V, F)
p = fp.octree_grid_setup(V, 1 << 20); p.c.stop_extent = 1 << E
for i in range(reps):
    ctx.sync(); t = time.perf_counter()
    o = fp.Octree.build(ctx, m, p)
    ctx.sync(); dt = time.perf_counter() - t
    print(f"build {i}: {dt*1e3:.2f} ms  {o.sizes()}  kernel-timer {ctx.last_kernel_ms():.2f} ms", flush=True)
    if i < reps - 1:
        o.close()
t = time.perf_counter(); Vh, H, _ = o.hexes(); print(f"hexes export {1e3*(time.perf_counter()-t):.2f} ms")
