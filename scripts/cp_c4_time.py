"""Times of the two C4 query sets on the C3 mesh, resident path (env switches are read by the library)."""
import sys, os, numpy as np
sys.path.insert(0, ".")
import torch
import fpohm_b200 as fp
ctx = fp.Context(0)
V, F = fp.procedural.c3_mesh()
mesh = fp.TriMesh(ctx, V, F); mesh.build_aabb_tree()
dev = torch.device("cuda", 0); st = torch.cuda.current_stream()
tot = 0
for name, P in zip(("project", "classify"), fp.procedural.c4_queries(V, F)):
    n = len(P)
    dP = torch.from_numpy(P).to(dev); dS = torch.empty(n, dtype=torch.float64, device=dev); dI = torch.empty(n, dtype=torch.int32, device=dev)
    dC = torch.empty(n, 3, dtype=torch.float64, device=dev); dN = torch.empty(n, 3, dtype=torch.float64, device=dev)
    f = lambda: mesh.signed_distance_dev(dP.data_ptr(), n, dS.data_ptr(), dI.data_ptr(), dC.data_ptr(), dN.data_ptr(), st.cuda_stream)
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(5): f()
    b.record(st); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5; tot += ms
    print(f"{os.environ.get('TAG','')} {name:9s} n={n:9d} {ms:8.3f} ms K1 {ctx.query_kernel_ms(5):.3f} ms", flush=True)
print(f"{os.environ.get('TAG','')} step {tot:.3f} ms = {11857634/tot/1e3:.1f} Mq/s")
