"""One C4 step (both query sets on the C3 mesh) through the resident path, argv[1] repetitions — for ncu captures."""
import sys, numpy as np
sys.path.insert(0, ".")
import torch
import fpohm_b200 as fp
ctx = fp.Context(0)
V, F = fp.procedural.c3_mesh()
mesh = fp.TriMesh(ctx, V, F); mesh.build_aabb_tree()
dev = torch.device("cuda", 0); st = torch.cuda.current_stream()
bufs = []
for P in fp.procedural.c4_queries(V, F):
    n = len(P)
    bufs.append((torch.from_numpy(P).to(dev), n, torch.empty(n, dtype=torch.float64, device=dev), torch.empty(n, dtype=torch.int32, device=dev),
                 torch.empty(n, 3, dtype=torch.float64, device=dev), torch.empty(n, 3, dtype=torch.float64, device=dev)))
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    for dP, n, dS, dI, dC, dN in bufs:
        mesh.signed_distance_dev(dP.data_ptr(), n, dS.data_ptr(), dI.data_ptr(), dC.data_ptr(), dN.data_ptr(), st.cuda_stream)
torch.cuda.synchronize()
