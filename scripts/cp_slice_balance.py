"""How evenly does the C4 job split over N ranks?  Times every rank's slice of both query sets on ONE GPU, for contiguous ranges and for
block-cyclic slices (M blocks per rank and part): max / mean over the ranks is the strong-scaling loss that comes from the split alone."""
import sys, numpy as np
sys.path.insert(0, ".")
import torch
import fpohm_b200 as fp
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ctx = fp.Context(0)
V, F = fp.procedural.c3_mesh()
mesh = fp.TriMesh(ctx, V, F); mesh.build_aabb_tree()
dev = torch.device("cuda", 0); st = torch.cuda.current_stream()
sets = fp.procedural.c4_queries(V, F)

def time_slice(P):
    n = len(P)
    dP = torch.from_numpy(np.ascontiguousarray(P)).to(dev); dS = torch.empty(n, dtype=torch.float64, device=dev); dI = torch.empty(n, dtype=torch.int32, device=dev)
    dC = torch.empty(n, 3, dtype=torch.float64, device=dev); dN = torch.empty(n, 3, dtype=torch.float64, device=dev)
    f = lambda: mesh.signed_distance_dev(dP.data_ptr(), n, dS.data_ptr(), dI.data_ptr(), dC.data_ptr(), dN.data_ptr(), st.cuda_stream)
    for _ in range(2): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(3): f()
    b.record(st); torch.cuda.synchronize()
    return a.elapsed_time(b) / 3

for M in ([int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else (1, 2, 4, 8)):
    per_rank = np.zeros(N)
    for P in sets:
        n = len(P); B = -(-n // (N * M))
        for r in range(N):
            idx = np.concatenate([np.arange(min((j * N + r) * B, n), min((j * N + r + 1) * B, n)) for j in range(M)])
            per_rank[r] += time_slice(P[idx])
    print(f"N={N} M={M} per-rank ms {np.round(per_rank, 2).tolist()} max {per_rank.max():.2f} mean {per_rank.mean():.2f} max/mean {per_rank.max() / per_rank.mean():.3f}", flush=True)
