"""Run only one query set through the resident query path a few times (for ncu captures).  argv: [gear|c3proj] [reps]"""
import sys, numpy as np
sys.path.insert(0, ".")
import torch
import fpohm_b200 as fp
import bench
which = sys.argv[1] if len(sys.argv) > 1 else "gear"
ctx = fp.Context(0)
if which == "gear":
    V, F, _ = fp.procedural.gear()
    mesh = fp.TriMesh(ctx, V, F)
    prm = fp.octree_grid_setup(V, 1 << 20); prm.c.stop_extent = 1 << 12
    mesh.build_aabb_tree()
    o = fp.Octree.build(ctx, mesh, prm)
    Vh, H, _ = o.hexes()
    ext = (Vh[H[:, 1].astype(np.int64), 0] - Vh[H[:, 0].astype(np.int64), 0])
    P = bench.make_queries(Vh, H, ext)
else:
    V, F = fp.procedural.c3_mesh()
    mesh = fp.TriMesh(ctx, V, F)
    mesh.build_aabb_tree()
    P, cls = fp.procedural.c4_queries(V, F)
    if which == "c3cls":
        P = cls
dev = torch.device("cuda", 0); st = torch.cuda.current_stream(); n = len(P)
dP = torch.from_numpy(P).to(dev)
dS = torch.empty(n, dtype=torch.float64, device=dev); dI = torch.empty(n, dtype=torch.int32, device=dev)
dC = torch.empty(n, 3, dtype=torch.float64, device=dev); dN = torch.empty(n, 3, dtype=torch.float64, device=dev)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    mesh.signed_distance_dev(dP.data_ptr(), n, dS.data_ptr(), dI.data_ptr(), dC.data_ptr(), dN.data_ptr(), st.cuda_stream)
torch.cuda.synchronize()
