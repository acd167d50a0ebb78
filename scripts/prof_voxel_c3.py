"""1024^3 z-ray parity voxelization of the C3 mesh, resident output (for an ncu launch list / timing): argv[1] = passes."""
import sys, time
sys.path.insert(0, ".")
import torch
import fpohm_b200 as fp
ctx = fp.Context(0)
V, F = fp.procedural.c3_mesh()
m = fp.TriMesh(ctx, V, F)
g = fp.VoxelGrid(V.min(0), V.max(0) - V.min(0), 1.0 / 1024, 0)
d = torch.empty(g.num_voxels(), dtype=torch.uint8, device="cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(2):
    fp.voxel_sign_dev(ctx, m, g, d.data_ptr(), 0)
ctx.sync(); torch.cuda.synchronize()
ts = []
for _ in range(n):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); fp.voxel_sign_dev(ctx, m, g, d.data_ptr(), 0); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("voxel_sign 1024^3 ms:", [round(t, 3) for t in ts], "dims", g.dims.tolist(), "filled", int(d.sum().item()))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    fp.voxel_sign_dev(ctx, m, g, d.data_ptr(), 0)
b.record(); torch.cuda.synchronize()
print("10 calls back to back (the bench's way): %.4f ms per call" % (a.elapsed_time(b) / 10))
