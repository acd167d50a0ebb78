"""Does Morton-sorting the queries help the packet search?  bench order vs Morton order vs shuffled (resident path)."""
import sys, numpy as np
sys.path.insert(0, ".")
import torch
import fpohm_b200 as fp
import bench
ctx = fp.Context(0)
V, F = bench.workload(fp)
mesh = fp.TriMesh(ctx, V, F)
prm = fp.octree_grid_setup(V, 1 << 20); prm.c.stop_extent = 1 << bench.STOP_E
mesh.build_aabb_tree()
o = fp.Octree.build(ctx, mesh, prm)
Vh, H, _ = o.hexes()
ext = (Vh[H[:, 1].astype(np.int64), 0] - Vh[H[:, 0].astype(np.int64), 0])
P = bench.make_queries(Vh, H, ext)
def morton(P, bits=16):
    mn, mx = P.min(0), P.max(0)
    q = ((P - mn) / (mx - mn).max() * ((1 << bits) - 1)).astype(np.uint64)
    def spread(x):
        r = np.zeros_like(x)
        for b in range(bits):
            r |= ((x >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b)
        return r
    return spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1)) | (spread(q[:, 2]) << np.uint64(2))
order = np.argsort(morton(P), kind="stable")
rng = np.random.default_rng(0)
sets = {"bench order": P, "morton order": P[order], "shuffled": P[rng.permutation(len(P))]}
dev = torch.device("cuda", 0); st = torch.cuda.current_stream()
for name, Q in sets.items():
    Q = np.ascontiguousarray(Q); n = len(Q)
    dP = torch.from_numpy(Q).to(dev)
    dS = torch.empty(n, dtype=torch.float64, device=dev); dI = torch.empty(n, dtype=torch.int32, device=dev)
    dC = torch.empty(n, 3, dtype=torch.float64, device=dev); dN = torch.empty(n, 3, dtype=torch.float64, device=dev)
    f = lambda: mesh.signed_distance_dev(dP.data_ptr(), n, dS.data_ptr(), dI.data_ptr(), dC.data_ptr(), dN.data_ptr(), st.cuda_stream)
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(5): f()
    b.record(st); torch.cuda.synchronize()
    print(f"{name:13s} {a.elapsed_time(b)/5:8.3f} ms  packet kernel {ctx.query_kernel_ms(5):.3f} ms", flush=True)
