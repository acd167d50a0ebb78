"""Query-structure build time (igl tree + normals + flattenings), device vs host tree builder."""
import sys, time, os
sys.path.insert(0, ".")
import numpy as np
import fpohm_b200 as fp
ctx = fp.Context(0)
for name, (V, F) in {"gear 200k": fp.procedural.gear()[:2], "C3 2.03M": fp.procedural.c3_mesh()}.items():
    for rep in range(3):
        m = fp.TriMesh(ctx, V, F)
        t = time.perf_counter(); m.build_aabb_tree(); ctx.sync(); dt = time.perf_counter() - t
        m.close()
    print(f"{os.environ.get('TAG','device')} {name}: build_aabb_tree {dt*1e3:.1f} ms", flush=True)
