"""1024^3 dense predicate occupancy of the C3 mesh (for an ncu launch list / timing): argv[1] = passes."""
import sys
sys.path.insert(0, ".")
import fpohm_b200 as fp
ctx = fp.Context(0)
V, F = fp.procedural.c3_mesh()
m = fp.TriMesh(ctx, V, F)
g = fp.VoxelGrid(V.min(0), V.max(0) - V.min(0), 1.0 / 1024, 0)
ts = []
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    occ = fp.voxel_occupancy(ctx, m, g); ts.append(ctx.last_kernel_ms())
print("occupancy 1024^3 kernel ms:", [round(t, 3) for t in ts], "set", int(occ.sum()))
