"""Query-structure build time (igl tree + normals + flattenings): gear, the C3 mesh (tied barycentres on all axes -> host std::sort per axis) and the
C3 mesh with its vertices jittered by 1e-9 (no ties: everything but acos on the device).  FPOHM_TREE_TIMELINE=1 prints the stages."""
import sys, time, os
sys.path.insert(0, ".")
import numpy as np
import fpohm_b200 as fp
ctx = fp.Context(0)
V3, F3 = fp.procedural.c3_mesh()
Vj = V3 + np.random.default_rng(5).uniform(-1e-9, 1e-9, V3.shape)
for name, (V, F) in {"gear 200k": fp.procedural.gear()[:2], "C3 2.03M": (V3, F3), "C3 2.03M jittered (tie-free)": (Vj, F3)}.items():
    for rep in range(3):
        m = fp.TriMesh(ctx, V, F)
        t = time.perf_counter(); m.build_aabb_tree(); ctx.sync(); dt = time.perf_counter() - t
        m.close()
    print(f"{os.environ.get('TAG','device')} {name}: build_aabb_tree {dt*1e3:.1f} ms", flush=True)
