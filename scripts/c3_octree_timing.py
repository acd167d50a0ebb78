import sys, time, os
if os.environ.get("WITH_TORCH"):
    import torch; torch.cuda.current_stream(); torch.zeros(1, device="cuda")
sys.path.insert(0, "/root/repo")
import numpy as np
import fpohm_b200 as fp
ctx = fp.Context(0)
pm = fp.procedural
V3, F3 = pm.midpoint_subdivide(*pm.linked_tori(4, 90, 44), 1)
mesh3 = fp.TriMesh(ctx, V3, F3)
p3 = fp.octree_grid_setup(V3, 1 << 20); p3.c.stop_extent = 1 << 10
for i in range(14):
    ctx.sync(); t0 = time.perf_counter(); o3 = fp.Octree.build(ctx, mesh3, p3); ctx.sync(); dt = (time.perf_counter() - t0) * 1e3
    t1 = time.perf_counter(); o3.close(); ctx.sync(); print(i, "build %.1f ms  close %.1f ms" % (dt, (time.perf_counter() - t1) * 1e3), flush=True)
