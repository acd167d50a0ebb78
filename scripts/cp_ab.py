"""A/B of the closest-point kernels (FPOHM_CP_MODE=0 per-lane igl order, 1 binary packets, 2 wide packets): timing + result
dump for bit comparison.
usage: FPOHM_CP_MODE=m python scripts/cp_ab.py out.npz [gear] [c3] ; python scripts/cp_ab.py --compare a.npz b.npz"""
import sys, os, time, numpy as np
sys.path.insert(0, ".")
if sys.argv[1] == "--compare":
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    ok = True
    for k in a.files:
        if k not in b.files:
            continue
        same = np.array_equal(a[k], b[k], equal_nan=True)
        ok &= same
        if not same:
            bad = np.flatnonzero((a[k] != b[k]).reshape(len(a[k]), -1).any(1))
            print("DIFF", k, len(bad), "rows, first", bad[:5])
    print("IDENTICAL" if ok else "MISMATCH")
    sys.exit(0 if ok else 1)
import torch
import fpohm_b200 as fp
import bench
which = set(sys.argv[2:]) or {"gear", "c3"}
ctx = fp.Context(0)
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream()
out = {}


def run(mesh, sets, tag):
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
        ms = a.elapsed_time(b) / 5
        print(f"mode={os.environ.get('FPOHM_CP_MODE','2')} sort={os.environ.get('FPOHM_CP_SORT','auto')} {tag}:{name:9s} n={n:8d} {ms:8.3f} ms  {n/ms/1e3:8.1f} Mq/s  "
              f"K1 {ctx.query_kernel_ms(5):.3f} ms", flush=True)
        k = f"{tag}_{name}"
        out[k + "_S"] = dS.cpu().numpy(); out[k + "_I"] = dI.cpu().numpy(); out[k + "_C"] = dC.cpu().numpy(); out[k + "_N"] = dN.cpu().numpy()


rng = np.random.default_rng(5)
if "gear" in which:
    V, F, _ = fp.procedural.gear()
    mesh = fp.TriMesh(ctx, V, F)
    prm = fp.octree_grid_setup(V, 1 << 20); prm.c.stop_extent = 1 << 12
    mesh.build_aabb_tree()
    o = fp.Octree.build(ctx, mesh, prm)
    Vh, H, _ = o.hexes()
    ext = (Vh[H[:, 1].astype(np.int64), 0] - Vh[H[:, 0].astype(np.int64), 0])
    P = bench.make_queries(Vh, H, ext)
    run(mesh, {"bench": P, "shuffled": P[rng.permutation(len(P))[:1 << 20]],
               "far": rng.uniform(V.min(0) - 2, V.max(0) + 2, (1 << 18, 3)),
               "onverts": np.repeat(V, 2, 0)[: 1 << 18], "hexverts": Vh[: 1 << 20]}, "gear")
    o.close(); mesh.close()
if "c3" in which:
    V, F = fp.procedural.c3_mesh()
    mesh = fp.TriMesh(ctx, V, F)
    t = time.perf_counter(); mesh.build_aabb_tree(); print(f"c3 tree build {time.perf_counter()-t:.2f} s", flush=True)
    proj, cls = fp.procedural.c4_queries(V, F)
    prm = fp.octree_grid_setup(V, 1 << 20); prm.c.stop_extent = 1 << 11
    o = fp.Octree.build(ctx, mesh, prm)
    Vh, H, _ = o.hexes()
    cen = np.ascontiguousarray(Vh[H.astype(np.int64)].mean(1))
    print("c3 e11 leaves", len(H), flush=True)
    run(mesh, {"project": proj, "classify": cls, "leaf_e11": cen, "onverts": V[: 1 << 19]}, "c3")
np.savez(sys.argv[1], **out)
