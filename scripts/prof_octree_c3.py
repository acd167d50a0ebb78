"""One warm 1024^3-equivalent octree build of the C3 mesh (for an ncu launch list): argv[1] = number of builds."""
import sys
sys.path.insert(0, ".")
import fpohm_b200 as fp
ctx = fp.Context(0)
V, F = fp.procedural.c3_mesh()
m = fp.TriMesh(ctx, V, F)
p = fp.octree_grid_setup(V, 1 << 20); p.c.stop_extent = 1 << 10
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    o = fp.Octree.build(ctx, m, p); ctx.sync(); o.close()
